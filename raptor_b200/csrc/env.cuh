// raptor_b200/csrc/env.cuh -- device-side quadrotor environment: one environment per thread, state in registers.
//
// What it computes follows the reference (file:line cited per function; L2F/ = rl_tools/rl/environments/l2f/);
// how it is organised does not: parameters are read through an accessor that resolves at compile time to shared
// memory (dynamics block, staged once per launch) or to the read-only global path, loop invariants of the four RK4
// dynamics evaluations (1/m, F_d/m, J^-1 tau_d, 1/tau) are hoisted out of the time loop, and the non-integrated
// parts of the state never leave registers between steps.
#pragma once
#include "layout.h"
#include "rng.cuh"

namespace b200l2f {

// ---------------------------------------------------------------------------------------------------------------
// parameter accessors
// ---------------------------------------------------------------------------------------------------------------
struct ParamsGlobal {  // struct-of-arrays in HBM: element i of environment e at base[i * stride + e]
    const float* __restrict__ base;  // already offset by the environment index
    size_t stride;
    __device__ __forceinline__ float operator[](int i) const { return __ldg(base + (size_t)i * stride); }
};
// dynamics block [0, P_DYN_DIM) staged in shared memory as sm[i * blockDim.x + tid]; the rest from HBM/L1.
// NC: the parameter buffer is read-only for the whole launch (ld.global.nc); kernels that rewrite parameters (collect's reset) use NC = false.
template <bool NC>
struct ParamsStagedT {
    const float* sm;   // already offset by threadIdx.x
    int sm_stride;
    const float* base;
    size_t stride;
    __device__ __forceinline__ float operator[](int i) const {
        if(i < P_DYN_DIM) return sm[i * sm_stride];
        return NC ? __ldg(base + (size_t)i * stride) : base[(size_t)i * stride];
    }
};
using ParamsStaged = ParamsStagedT<true>;

__device__ __forceinline__ float clampf(float x, float lo, float hi){ return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------------------------------------------
// per-thread state.  x = integrated fields: position[3] orientation[4] linear_velocity[3] angular_velocity[3] rpm[4]
// (the REQUIRES_INTEGRATION layers, L2F/operations_generic/50_state_algebra.h:22-53)
// ---------------------------------------------------------------------------------------------------------------
enum { X_POS = 0, X_ORI = 3, X_VEL = 7, X_OMEGA = 10, X_RPM = 13, X_DIM = 17 };

template <class Spec>
struct EnvState {
    float x[X_DIM];
    float last_action[4];
    float force[3], torque[3];
    int current_step;
    int traj_type;
    float lang[Spec::LANGEVIN ? 12 : 1];  // position[3] velocity[3] position_raw[3] velocity_raw[3]
    float hist[Spec::H == 1 ? 4 : 1];     // H == 1: the single history slot lives in registers; H > 1: it stays in HBM (hist_ptr)
};

// SoA state buffer S[STATE_DIM][n]; `s` already offset by the environment index
template <class Spec>
__device__ __forceinline__ void load_state(EnvState<Spec>& st, const float* __restrict__ s, size_t n){
#pragma unroll
    for(int i = 0; i < 13; i++) st.x[i] = s[(size_t)i * n];
#pragma unroll
    for(int i = 0; i < 4; i++) st.x[X_RPM + i] = s[(size_t)(S_RPM + i) * n];
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = s[(size_t)(S_LAST_ACTION + i) * n];
#pragma unroll
    for(int i = 0; i < 3; i++){ st.force[i] = s[(size_t)(S_FORCE + i) * n]; st.torque[i] = s[(size_t)(S_TORQUE + i) * n]; }
    st.current_step = (int)s[(size_t)S_CURRENT_STEP * n];
    st.traj_type = (int)s[(size_t)s_traj_type(Spec::H) * n];
    if constexpr(Spec::LANGEVIN){
#pragma unroll
        for(int i = 0; i < 12; i++) st.lang[i] = s[(size_t)(s_langevin(Spec::H) + i) * n];
    }
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) st.hist[i] = s[(size_t)(S_HIST + i) * n];
    }
}
// same through L2 only (ld.global.cg): used where another CTA of the SAME launch may have written the buffer (time-chunked scheduler)
template <class Spec>
__device__ __forceinline__ void load_state_cg(EnvState<Spec>& st, const float* s, size_t n){
#pragma unroll
    for(int i = 0; i < 13; i++) st.x[i] = __ldcg(s + (size_t)i * n);
#pragma unroll
    for(int i = 0; i < 4; i++) st.x[X_RPM + i] = __ldcg(s + (size_t)(S_RPM + i) * n);
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = __ldcg(s + (size_t)(S_LAST_ACTION + i) * n);
#pragma unroll
    for(int i = 0; i < 3; i++){ st.force[i] = __ldcg(s + (size_t)(S_FORCE + i) * n); st.torque[i] = __ldcg(s + (size_t)(S_TORQUE + i) * n); }
    st.current_step = (int)__ldcg(s + (size_t)S_CURRENT_STEP * n);
    st.traj_type = (int)__ldcg(s + (size_t)s_traj_type(Spec::H) * n);
    if constexpr(Spec::LANGEVIN){
#pragma unroll
        for(int i = 0; i < 12; i++) st.lang[i] = __ldcg(s + (size_t)(s_langevin(Spec::H) + i) * n);
    }
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) st.hist[i] = __ldcg(s + (size_t)(S_HIST + i) * n);
    }
}
// stores everything except the H > 1 action-history ring (maintained in place by the callers)
template <class Spec>
__device__ __forceinline__ void store_state(const EnvState<Spec>& st, float* __restrict__ s, size_t n){
#pragma unroll
    for(int i = 0; i < 13; i++) s[(size_t)i * n] = st.x[i];
#pragma unroll
    for(int i = 0; i < 4; i++) s[(size_t)(S_RPM + i) * n] = st.x[X_RPM + i];
#pragma unroll
    for(int i = 0; i < 4; i++) s[(size_t)(S_LAST_ACTION + i) * n] = st.last_action[i];
#pragma unroll
    for(int i = 0; i < 3; i++) s[(size_t)(S_ANGVEL_HIST + i) * n] = st.x[X_OMEGA + i];  // history length 0: copy of omega (70_post_integration.h:63-67)
#pragma unroll
    for(int i = 0; i < 3; i++){ s[(size_t)(S_FORCE + i) * n] = st.force[i]; s[(size_t)(S_TORQUE + i) * n] = st.torque[i]; }
    s[(size_t)S_CURRENT_STEP * n] = (float)st.current_step;
    s[(size_t)s_traj_type(Spec::H) * n] = (float)st.traj_type;
    if constexpr(Spec::LANGEVIN){
#pragma unroll
        for(int i = 0; i < 12; i++) s[(size_t)(s_langevin(Spec::H) + i) * n] = st.lang[i];
    }
    else{
#pragma unroll
        for(int i = 0; i < 12; i++) s[(size_t)(s_langevin(Spec::H) + i) * n] = 0.0f;
    }
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) s[(size_t)(S_HIST + i) * n] = st.hist[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// invariants of the dynamics that do not change between resets
// ---------------------------------------------------------------------------------------------------------------
struct DynInvariants {
    float inv_mass;      // 1 / mass                         (60_dynamics.h:55)
    float fa[3];         // force / mass                     (60_dynamics.h:91-93)
    float ga[3];         // gravity + force / mass           (packed dynamics of the default-math kernels: one addend instead of two)
    float ta[3];         // J_inv * torque_disturbance       (60_dynamics.h:97)
    float half_range;    // (action_limit.max - min) / 2     (operations_generic.h:105)
    float dt;
};
template <class Spec, class P, bool FAST = false>
__device__ __forceinline__ void dyn_invariants(DynInvariants& d, const P& p, const EnvState<Spec>& st){
    const float mass = p[P_MASS];
    d.inv_mass = FAST ? rcp_approx(mass) : 1.0f / mass;
#pragma unroll
    for(int i = 0; i < 3; i++){ d.fa[i] = FAST ? st.force[i] * d.inv_mass : st.force[i] / mass; d.ga[i] = p[P_GRAVITY + i] + d.fa[i]; }
#pragma unroll
    for(int i = 0; i < 3; i++){
        float a = 0.0f;
#pragma unroll
        for(int j = 0; j < 3; j++) a += p[P_JINV + 3 * i + j] * st.torque[j];
        d.ta[i] = a;
    }
    d.half_range = (p[P_ACT_MAX] - p[P_ACT_MIN]) / 2.0f;
    d.dt = p[P_DT];
}

// ---------------------------------------------------------------------------------------------------------------
// multirotor_dynamics for the full state stack  (L2F/operations_generic/60_dynamics.h:18-72 base, :87-99 random force,
// :100-111 rotors; helpers L2F/quaternion_helper.h:11-18,22-35).  The body is driven by the CURRENT rpm, the setpoint only
// drives d(rpm)/dt (:102).  INV_TAU: the staged copies of the rotor time constants hold their reciprocals.
// ---------------------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ void dynamics(const P& p, const DynInvariants& d, const float* __restrict__ x, const float* __restrict__ setpoint, float* __restrict__ dx){
    float thrust[3] = {0.0f, 0.0f, 0.0f}, torque[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for(int r = 0; r < 4; r++){
        const float rpm = x[X_RPM + r];
        const float tm = p[P_THRUST_COEF + 3 * r + 0] + p[P_THRUST_COEF + 3 * r + 1] * rpm + p[P_THRUST_COEF + 3 * r + 2] * rpm * rpm;
        float rt[3];
#pragma unroll
        for(int i = 0; i < 3; i++) rt[i] = p[P_THRUST_DIR + 3 * r + i] * tm;
#pragma unroll
        for(int i = 0; i < 3; i++) thrust[i] += rt[i];
        const float sc = tm * p[P_TORQUE_CONST + r];
#pragma unroll
        for(int i = 0; i < 3; i++) torque[i] += p[P_TORQUE_DIR + 3 * r + i] * sc;
        const float px = p[P_ROTOR_POS + 3 * r + 0], py = p[P_ROTOR_POS + 3 * r + 1], pz = p[P_ROTOR_POS + 3 * r + 2];
        torque[0] += py * rt[2] - pz * rt[1];
        torque[1] += pz * rt[0] - px * rt[2];
        torque[2] += px * rt[1] - py * rt[0];
    }
#pragma unroll
    for(int i = 0; i < 3; i++) dx[X_POS + i] = x[X_VEL + i];
    const float q0 = x[X_ORI], q1 = x[X_ORI + 1], q2 = x[X_ORI + 2], q3 = x[X_ORI + 3];
    const float w0 = x[X_OMEGA], w1 = x[X_OMEGA + 1], w2 = x[X_OMEGA + 2];
    dx[X_ORI + 0] = (-q1 * w0 - q2 * w1 - q3 * w2) * 0.5f;
    dx[X_ORI + 1] = ( q0 * w0 + q2 * w2 - q3 * w1) * 0.5f;
    dx[X_ORI + 2] = ( q0 * w1 + q3 * w0 - q1 * w2) * 0.5f;
    dx[X_ORI + 3] = ( q0 * w2 + q1 * w1 - q2 * w0) * 0.5f;
    {   // v + q0 (2 q x v) + q x (2 q x v), then /m, + g, + F_d/m
        float v0 = (q2 * thrust[2] - q3 * thrust[1]) * 2.0f;
        float v1 = (q3 * thrust[0] - q1 * thrust[2]) * 2.0f;
        float v2 = (q1 * thrust[1] - q2 * thrust[0]) * 2.0f;
        float o0 = q2 * v2 - q3 * v1;
        float o1 = q3 * v0 - q1 * v2;
        float o2 = q1 * v1 - q2 * v0;
        o0 += v0 * q0; o1 += v1 * q0; o2 += v2 * q0;
        o0 += thrust[0]; o1 += thrust[1]; o2 += thrust[2];
        dx[X_VEL + 0] = o0 * d.inv_mass + p[P_GRAVITY + 0] + d.fa[0];
        dx[X_VEL + 1] = o1 * d.inv_mass + p[P_GRAVITY + 1] + d.fa[1];
        dx[X_VEL + 2] = o2 * d.inv_mass + p[P_GRAVITY + 2] + d.fa[2];
    }
    {   // J^-1 (tau - w x J w) + J^-1 tau_d
        float v[3];
#pragma unroll
        for(int i = 0; i < 3; i++) v[i] = p[P_J + 3 * i + 0] * w0 + p[P_J + 3 * i + 1] * w1 + p[P_J + 3 * i + 2] * w2;
        const float c0 = w1 * v[2] - w2 * v[1];
        const float c1 = w2 * v[0] - w0 * v[2];
        const float c2 = w0 * v[1] - w1 * v[0];
        const float t0 = torque[0] - c0, t1 = torque[1] - c1, t2 = torque[2] - c2;
#pragma unroll
        for(int i = 0; i < 3; i++) dx[X_OMEGA + i] = p[P_JINV + 3 * i + 0] * t0 + p[P_JINV + 3 * i + 1] * t1 + p[P_JINV + 3 * i + 2] * t2 + d.ta[i];
    }
#pragma unroll
    for(int r = 0; r < 4; r++){  // first-order motor lag with separate rising / falling time constants (reciprocals staged)
        const float rpm = x[X_RPM + r];
        const float inv_tau = setpoint[r] >= rpm ? p[P_TAU_RISE + r] : p[P_TAU_FALL + r];
        dx[X_RPM + r] = (setpoint[r] - rpm) * inv_tau;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// step = action scaling + RK4 + post_integration.
//   rl_tools::step  L2F/operations_generic.h:94-130;  rk4  rl_tools/utils/generic/integrators.h:18-50;
//   post_integration  L2F/operations_generic/70_post_integration.h:20-37 (normalise, clamp), :40-48 (last action),
//   :59-81 (omega history), :85-100 (rpm clamp), :112-125 (action ring buffer), :127-170 (Langevin target).
// `p` must be an accessor whose time-constant slots hold RECIPROCALS (see stage_dynamics / ParamsInvTau).
// hist_ptr: H > 1 only, the SoA rows of action_history of the NEXT state (element (h,a) at hist_ptr[(4h+a)*n]).
// ---------------------------------------------------------------------------------------------------------------
template <class Spec, bool NOISE, class P, bool ROLLED_RK4 = false>
__device__ __forceinline__ void env_step(EnvState<Spec>& st, const P& p, const DynInvariants& d, const float* __restrict__ action, uint64_t& rng,
                                         float* __restrict__ hist_ptr, size_t n){
    float setpoint[4];
#pragma unroll
    for(int i = 0; i < 4; i++){
        float a = action[i];
        if constexpr(NOISE) a += rng_normal_t<Spec::RNG_OOL>(rng, 0.0f, p[P_ACTION_NOISE]);
        a = clampf(a, -1.0f, 1.0f);
        setpoint[i] = a * d.half_range + p[P_ACT_MIN] + d.half_range;
    }
    const float dt = d.dt;
    const float dt2 = dt / 2.0f, dt3 = dt / 3.0f, dt6 = dt / 6.0f;
    float k[X_DIM], tmp[X_DIM], acc[X_DIM];
    if constexpr(ROLLED_RK4){
        // same arithmetic as below with the four stages as ONE loop body (a quarter of the code; the stage weights are selected per iteration)
#pragma unroll
        for(int i = 0; i < X_DIM; i++){ acc[i] = st.x[i]; tmp[i] = st.x[i]; }
#pragma unroll 1
        for(int s = 0; s < 4; s++){
            dynamics(p, d, tmp, setpoint, k);
            const float wa = (s == 0 || s == 3) ? dt6 : dt3;     // next = x + dt/6 k1 + dt/3 k2 + dt/3 k3 + dt/6 k4, accumulated in that order
            const float wt = (s == 2) ? dt : dt2;                 // stage inputs x + dt/2 k1, x + dt/2 k2, x + dt k3
#pragma unroll
            for(int i = 0; i < X_DIM; i++){ acc[i] += wa * k[i]; tmp[i] = st.x[i] + wt * k[i]; }
        }
#pragma unroll
        for(int i = 0; i < X_DIM; i++) st.x[i] = acc[i];
    }
    else{
        dynamics(p, d, st.x, setpoint, k);                                            // k1
#pragma unroll
        for(int i = 0; i < X_DIM; i++){ acc[i] = st.x[i] + dt6 * k[i]; tmp[i] = st.x[i] + dt2 * k[i]; }
        dynamics(p, d, tmp, setpoint, k);                                             // k2
#pragma unroll
        for(int i = 0; i < X_DIM; i++){ acc[i] += dt3 * k[i]; tmp[i] = st.x[i] + dt2 * k[i]; }
        dynamics(p, d, tmp, setpoint, k);                                             // k3
#pragma unroll
        for(int i = 0; i < X_DIM; i++){ acc[i] += dt3 * k[i]; tmp[i] = st.x[i] + dt * k[i]; }
        dynamics(p, d, tmp, setpoint, k);                                             // k4
#pragma unroll
        for(int i = 0; i < X_DIM; i++) st.x[i] = acc[i] + dt6 * k[i];
    }
    // ---- post integration
    {
        float nrm = 0.0f;
#pragma unroll
        for(int i = 0; i < 4; i++) nrm += st.x[X_ORI + i] * st.x[X_ORI + i];
        nrm = sqrtf(nrm);
#pragma unroll
        for(int i = 0; i < 4; i++) st.x[X_ORI + i] = st.x[X_ORI + i] / nrm;
#pragma unroll
        for(int i = 0; i < 3; i++){
            st.x[X_POS + i] = clampf(st.x[X_POS + i], -100000.0f, 100000.0f);
            st.x[X_VEL + i] = clampf(st.x[X_VEL + i], -100000.0f, 100000.0f);
            st.x[X_OMEGA + i] = clampf(st.x[X_OMEGA + i], -100000.0f, 100000.0f);
        }
    }
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = action[i];
#pragma unroll
    for(int i = 0; i < 4; i++) st.x[X_RPM + i] = clampf(st.x[X_RPM + i], p[P_ACT_MIN], p[P_ACT_MAX]);
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) st.hist[i] = action[i];   // current_step stays 0: (0 + 1) % 1
    }
    else{
        const int cs = st.current_step;
#pragma unroll
        for(int i = 0; i < 4; i++) hist_ptr[(size_t)(4 * cs + i) * n] = action[i];
        st.current_step = (cs + 1) % Spec::H;
    }
    if constexpr(Spec::LANGEVIN){
        if(st.traj_type == 1){
            const float gamma = p[P_LANGEVIN_GAMMA], omega = p[P_LANGEVIN_OMEGA], sigma = p[P_LANGEVIN_SIGMA], alpha = p[P_LANGEVIN_ALPHA];
            const float sqrt_dt = sqrtf(dt);
#pragma unroll
            for(int dim = 0; dim < 3; dim++){
                const float x_prev = st.lang[6 + dim];
                const float v_prev = st.lang[9 + dim];
                const float dW = sqrt_dt * rng_normal_t<Spec::RNG_OOL>(rng, 0.0f, 1.0f);
                const float v_next = v_prev + (-gamma * v_prev - omega * omega * x_prev) * dt + sigma * dW;
                const float x_next = x_prev + v_next * dt;
                st.lang[6 + dim] = x_next;
                st.lang[9 + dim] = v_next;
                const float v_smooth = alpha * v_next + (1.0f - alpha) * st.lang[3 + dim];
                const float x_smooth = st.lang[dim] + v_smooth * dt;
                st.lang[dim] = x_smooth;
                st.lang[3 + dim] = v_smooth;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// desired state (L2F/operations_generic/35_get_desired_state.h:17-50)
// ---------------------------------------------------------------------------------------------------------------
template <class Spec>
__device__ __forceinline__ void desired_state(const EnvState<Spec>& st, float dpos[3], float dvel[3]){
#pragma unroll
    for(int i = 0; i < 3; i++){ dpos[i] = 0.0f; dvel[i] = 0.0f; }
    if constexpr(Spec::LANGEVIN){
        if(st.traj_type == 1){
#pragma unroll
            for(int i = 0; i < 3; i++){ dpos[i] = st.lang[i]; dvel[i] = st.lang[3 + i]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// observe: first 18 columns + most recent action (the 22 columns every actor on the path consumes first).
//   L2F/operations_generic/40_observe.h: Position :42-60 / TrajectoryTrackingPosition :387-407, RotationMatrix :81-106,
//   LinearVelocity :108-125 / TrajectoryTrackingLinearVelocity :409-429, AngularVelocityDelayed<0> :221-254.
// Noise draws happen in column order (3 + 9 + 3 + 3 normals), each skipped when its std is 0.
// ---------------------------------------------------------------------------------------------------------------
// FAST: the noise draws use the MUFU Box-Muller of the default-math kernels (same integer stream, |error| ~5e-7 x std)
template <class Spec, bool NOISE, bool FAST = false, class P>
__device__ __forceinline__ void observe18(const EnvState<Spec>& st, const P& p, uint64_t& rng, float* __restrict__ o){
    float dpos[3], dvel[3];
    desired_state(st, dpos, dvel);
#pragma unroll
    for(int i = 0; i < 3; i++){
        float v = (Spec::OBS_LAYOUT == OBS_RAPTOR) ? st.x[X_POS + i] : st.x[X_POS + i] - dpos[i];
        if constexpr(NOISE) v += rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, p[P_NOISE_POS]);
        o[i] = v;
    }
    const float q0 = st.x[X_ORI], q1 = st.x[X_ORI + 1], q2 = st.x[X_ORI + 2], q3 = st.x[X_ORI + 3];
    o[3]  = (1 - 2 * q2 * q2 - 2 * q3 * q3);
    o[4]  = (    2 * q1 * q2 - 2 * q0 * q3);
    o[5]  = (    2 * q1 * q3 + 2 * q0 * q2);
    o[6]  = (    2 * q1 * q2 + 2 * q0 * q3);
    o[7]  = (1 - 2 * q1 * q1 - 2 * q3 * q3);
    o[8]  = (    2 * q2 * q3 - 2 * q0 * q1);
    o[9]  = (    2 * q1 * q3 - 2 * q0 * q2);
    o[10] = (    2 * q2 * q3 + 2 * q0 * q1);
    o[11] = (1 - 2 * q1 * q1 - 2 * q2 * q2);
    if constexpr(NOISE){
#pragma unroll
        for(int i = 0; i < 9; i++) o[3 + i] += rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, p[P_NOISE_ORI]);
    }
#pragma unroll
    for(int i = 0; i < 3; i++){
        float v = (Spec::OBS_LAYOUT == OBS_RAPTOR) ? st.x[X_VEL + i] : st.x[X_VEL + i] - dvel[i];
        if constexpr(NOISE) v += rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, p[P_NOISE_LINVEL]);
        o[12 + i] = v;
    }
#pragma unroll
    for(int i = 0; i < 3; i++){
        float v = st.x[X_OMEGA + i];
        if constexpr(NOISE) v += rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, p[P_NOISE_ANGVEL]);
        o[15 + i] = v;
    }
}
// action history, most recent first (40_observe.h:270-291) + rotor speeds for the teacher layout (:256-268).
// hist_ptr as in env_step (H > 1), ignored for H == 1.  o points at column 18.
template <class Spec, class P>
__device__ __forceinline__ void observe_tail(const EnvState<Spec>& st, const P& p, const float* __restrict__ hist_ptr, size_t n, float* __restrict__ o, int n_hist){
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) o[i] = st.hist[i];
    }
    else{
        int cur = st.current_step == 0 ? Spec::H - 1 : st.current_step - 1;
        for(int h = 0; h < n_hist; h++){
#pragma unroll
            for(int i = 0; i < 4; i++) o[4 * h + i] = hist_ptr[(size_t)(4 * cur + i) * n];
            cur = cur == 0 ? Spec::H - 1 : cur - 1;
        }
    }
    if constexpr(Spec::OBS_LAYOUT == OBS_TEACHER){
#pragma unroll
        for(int i = 0; i < 4; i++) o[4 * Spec::H + i] = (st.x[X_RPM + i] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2.0f - 1.0f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// terminated (L2F/operations_generic.h:142-166) and the Squared reward
// (L2F/parameters/reward_functions/squared/operations_generic.h:13-46, 48-58, 100-129).
// `prev` = state before the step (position/orientation/velocity/omega/last_action/desired state), `next` = after.
// ---------------------------------------------------------------------------------------------------------------
template <class P>
__device__ __forceinline__ bool env_terminated(const P& p, const float* __restrict__ x){
    bool t = false;
    if(p[P_TERM_ENABLED] != 0.0f){
        const float tp = p[P_TERM_POS], tv = p[P_TERM_LINVEL], tw = p[P_TERM_ANGVEL];
        // any |component| above its threshold == the three-input maximum of the |components| above it (NaN components compare false either way)
        t = max3(fabsf(x[X_POS]), fabsf(x[X_POS + 1]), fabsf(x[X_POS + 2])) > tp || max3(fabsf(x[X_VEL]), fabsf(x[X_VEL + 1]), fabsf(x[X_VEL + 2])) > tv ||
            max3(fabsf(x[X_OMEGA]), fabsf(x[X_OMEGA + 1]), fabsf(x[X_OMEGA + 2])) > tw;
    }
    return t;
}
struct RewardInputs {  // the slice of the previous state the reward needs (kept while the state is advanced in place)
    float pos[3], q3, vel[3], omega[3], last_action[4], dpos[3], dvel[3];
};
template <class Spec>
__device__ __forceinline__ void reward_inputs(RewardInputs& r, const EnvState<Spec>& st){
#pragma unroll
    for(int i = 0; i < 3; i++){ r.pos[i] = st.x[X_POS + i]; r.vel[i] = st.x[X_VEL + i]; r.omega[i] = st.x[X_OMEGA + i]; }
    r.q3 = st.x[X_ORI + 3];
#pragma unroll
    for(int i = 0; i < 4; i++) r.last_action[i] = st.last_action[i];
    desired_state(st, r.dpos, r.dvel);
}
// FAST (default-math fused kernels): a term whose weight is zero is skipped instead of multiplied by zero (same value for finite states; the
// foundation-policy reward has five zero weights out of eight), square roots on the MUFU unit.
template <bool FAST = false, class P>
__device__ __forceinline__ float env_reward(const P& p, const RewardInputs& s, const float* __restrict__ action, const float* __restrict__ xn, bool terminated_next, float dt){
    float weighted = 0.0f;
    auto on = [](float w){ return !FAST || w != 0.0f; };
    if(const float w = p[P_RW_POSITION]; on(w)){
        const float x = s.pos[0] - s.dpos[0], y = s.pos[1] - s.dpos[1], z = s.pos[2] - s.dpos[2];
        float c = sqrt_t<FAST>(x * x + y * y + z * z);
        const float clip = p[P_RW_POSITION_CLIP];
        if(clip > 0.0f) c = fminf(c, clip);
        weighted += w * c;
    }
    if(const float w = p[P_RW_ORIENTATION]; on(w)) weighted += w * (2.0f * (FAST ? acos_1m_fast(fabsf(s.q3)) : acosf(1.0f - fabsf(s.q3))));
    if(const float w = p[P_RW_LINVEL]; on(w)){
        const float x = s.vel[0] - s.dvel[0], y = s.vel[1] - s.dvel[1], z = s.vel[2] - s.dvel[2];
        weighted += w * sqrt_t<FAST>(x * x + y * y + z * z);
    }
    if(const float w = p[P_RW_ANGVEL]; on(w)) weighted += w * sqrt_t<FAST>(s.omega[0] * s.omega[0] + s.omega[1] * s.omega[1] + s.omega[2] * s.omega[2]);
    if(const float w = p[P_RW_LINACC]; on(w)){
        const float x = xn[X_VEL] - s.vel[0], y = xn[X_VEL + 1] - s.vel[1], z = xn[X_VEL + 2] - s.vel[2];
        weighted += w * (sqrt_t<FAST>(x * x + y * y + z * z) / dt);
    }
    if(const float w = p[P_RW_ANGACC]; on(w)){
        const float x = xn[X_OMEGA] - s.omega[0], y = xn[X_OMEGA + 1] - s.omega[1], z = xn[X_OMEGA + 2] - s.omega[2];
        weighted += w * (sqrt_t<FAST>(x * x + y * y + z * z) / dt);
    }
    if(const float w = p[P_RW_ACTION]; on(w)){
        const float hover = p[P_HOVER];
        float acc = 0.0f;
#pragma unroll
        for(int i = 0; i < 4; i++){ const float dd = (action[i] + 1.0f) / 2.0f - hover; acc += dd * dd; }
        float c = sqrt_t<FAST>(acc);
        weighted += w * (c * c);
    }
    if(const float w = p[P_RW_DACTION]; on(w)){
        float acc = 0.0f;
#pragma unroll
        for(int i = 0; i < 4; i++){ const float dd = action[i] - s.last_action[i]; acc += dd * dd; }
        weighted += w * sqrt_t<FAST>(acc);
    }
    const float scaled = p[P_RW_SCALE] * weighted;
    float r;
    if(terminated_next){ r = p[P_RW_TERM_PENALTY]; }
    else{
        r = -scaled + p[P_RW_CONSTANT];
        r = (r > 0.0f || !(p[P_RW_NONNEG] != 0.0f)) ? r : 0.0f;
    }
    return r;
}

}  // namespace b200l2f
