// raptor_b200/csrc/samplers.cuh -- on-device parameter / initial-state samplers (one environment per thread).
//
//   sample_initial_parameters   L2F/operations_generic/10_sample_initial_parameters.h:20-23 (copy of env.parameters),
//                               :35-201 (domain randomisation), :30-34 (size-deviation factor: normal(mean=-range, std=range), sic)
//   initial_state               L2F/operations_generic/20_initial_state.h:21-121
//   sample_initial_state        L2F/operations_generic/30_sample_initial_state.h:21-40 (orientation), :42-85 (base), :124-152 (random
//                               force/torque), :161-183 (rotor speeds), :185-196 (action history), :198-228 (trajectory mixture)
// The RNG draw ORDER is part of the contract and is the reference's.  Float expression trees keep the reference's association and its
// float/double literal promotions; arithmetic that must not be contracted into FMAs uses the __f*_rn intrinsics so that the results are
// bit-identical to the oracle up to the last ulp of cbrtf/sinf/cosf/logf.
#pragma once
#include "layout.h"
#include "rng.cuh"
#include "env.cuh"

namespace b200l2f {

struct ParamsRW {  // read/write view of one environment's parameter column in the SoA buffer
    float* base;   // already offset by the environment index
    size_t stride;
    __device__ __forceinline__ float& operator[](int i) const { return base[(size_t)i * stride]; }
};

#define B200_MUL(a, b) __fmul_rn((a), (b))
#define B200_ADD(a, b) __fadd_rn((a), (b))
#define B200_SUB(a, b) __fsub_rn((a), (b))
#define B200_DIV(a, b) __fdiv_rn((a), (b))

// returns false if the DR ranges violate the reference's assert_exit conditions (the caller records the error)
template <bool DR>
__device__ __forceinline__ bool sample_parameters(const float* __restrict__ env_p, const ParamsRW& p, uint64_t& rng){
    for(int i = 0; i < PARAMS_DIM; i++) p[i] = env_p[i];
    if constexpr(!DR){ return true; }
    else{
        float t2w_nominal;
        float gravity_norm = sqrtf(B200_ADD(B200_ADD(B200_MUL(p[P_GRAVITY], p[P_GRAVITY]), B200_MUL(p[P_GRAVITY + 1], p[P_GRAVITY + 1])), B200_MUL(p[P_GRAVITY + 2], p[P_GRAVITY + 2])));
        {
            const float max_action = p[P_ACT_MAX];
            float max_thrust_nominal = 0.0f;
            for(int r = 0; r < 4; r++){
                float v = B200_ADD(B200_ADD(p[P_THRUST_COEF + 3 * r], B200_MUL(p[P_THRUST_COEF + 3 * r + 1], max_action)), B200_MUL(B200_MUL(p[P_THRUST_COEF + 3 * r + 2], max_action), max_action));
                max_thrust_nominal = B200_ADD(max_thrust_nominal, v);
            }
            t2w_nominal = B200_DIV(max_thrust_nominal, B200_MUL(p[P_MASS], gravity_norm));
        }
        if(!(p[P_DR_T2W_MIN] < p[P_DR_T2W_MAX]) || !(p[P_DR_T2W_MIN] >= 1.5f)) return false;
        const float t2w = rng_uniform(rng, p[P_DR_T2W_MIN], p[P_DR_T2W_MAX]);
        const float factor_t2w = B200_DIV(t2w, t2w_nominal);
        if(!(p[P_DR_MASS_MIN] < p[P_DR_MASS_MAX])) return false;
        const float size_min = cbrtf(p[P_DR_MASS_MIN]);
        const float size_max = cbrtf(p[P_DR_MASS_MAX]);
        const float size_new = rng_uniform(rng, size_min, size_max);
        float mass_new = B200_MUL(B200_MUL(size_new, size_new), size_new);
        mass_new = clampf(mass_new, p[P_DR_MASS_MIN], p[P_DR_MASS_MAX]);
        const float scale_relative = cbrtf(B200_DIV(mass_new, p[P_MASS]));
        const float factor_mass = B200_DIV(mass_new, p[P_MASS]);
        p[P_MASS] = mass_new;
        const float factor_coef = B200_MUL(factor_t2w, factor_mass);
        for(int i = 0; i < 12; i++) p[P_THRUST_COEF + i] = B200_MUL(p[P_THRUST_COEF + i], factor_coef);
        float t2i_factor;
        {
            const float max_thrust = B200_DIV(B200_MUL(B200_MUL(t2w, p[P_MASS]), gravity_norm), 4.0f);
            const float first_rotor_distance = fabsf(p[P_ROTOR_POS]);
            const float max_torque = (float)((double)first_rotor_distance * 1.414213562373095 * (double)max_thrust);
            const float t2i_nominal = B200_DIV(max_torque, p[P_J]);
            if(!(p[P_DR_T2I_MIN] < p[P_DR_T2I_MAX])) return false;
            const float t2i = rng_uniform(rng, p[P_DR_T2I_MIN], p[P_DR_T2I_MAX]);
            t2i_factor = B200_DIV(t2i, t2i_nominal);
        }
        if(p[P_DR_MASS_SIZE_DEV] == 0.0f) return false;
        float size_factor;
        {
            const float range = p[P_DR_MASS_SIZE_DEV];
            const float f = rng_normal(rng, -range, range);
            size_factor = f < 0.0f ? B200_DIV(1.0f, B200_SUB(1.0f, f)) : B200_ADD(1.0f, f);
        }
        const float rotor_distance_factor = B200_MUL(scale_relative, size_factor);
        {
            const float inertia_factor = B200_DIV(t2i_factor, rotor_distance_factor);
            for(int a = 0; a < 3; a++){
                p[P_J + 4 * a] = B200_DIV(p[P_J + 4 * a], inertia_factor);
                p[P_JINV + 4 * a] = B200_MUL(p[P_JINV + 4 * a], inertia_factor);
            }
            for(int i = 0; i < 12; i++) p[P_ROTOR_POS + i] = B200_MUL(p[P_ROTOR_POS + i], rotor_distance_factor);
            float max_rotor_distance = 0.0f;
            for(int r = 0; r < 4; r++){
                const float x = p[P_ROTOR_POS + 3 * r], y = p[P_ROTOR_POS + 3 * r + 1], z = p[P_ROTOR_POS + 3 * r + 2];
                const float dd = sqrtf(B200_ADD(B200_ADD(B200_MUL(x, x), B200_MUL(y, y)), B200_MUL(z, z)));
                if(dd > max_rotor_distance) max_rotor_distance = dd;
            }
            p[P_TERM_POS] = B200_MUL(max_rotor_distance, 20.0f);
            p[P_INIT_MAX_POS] = B200_MUL(max_rotor_distance, 10.0f);
        }
        if(p[P_DR_KQ_MIN] == 0.0f || p[P_DR_KQ_MAX] == 0.0f) return false;
        {
            const float kq = rng_uniform(rng, p[P_DR_KQ_MIN], p[P_DR_KQ_MAX]);
            for(int r = 0; r < 4; r++) p[P_TORQUE_CONST + r] = kq;
        }
        if(p[P_DR_DIST_FORCE_MAX] == 0.0f) return false;
        {
            float surplus = (float)((double)t2w - 1.0);
            if(surplus < 0.0f) surplus = 0.0f;
            const float multiple = rng_uniform(rng, 0.0f, B200_MUL(surplus, p[P_DR_DIST_FORCE_MAX]));
            p[P_DIST_FORCE_MEAN] = 0.0f;
            p[P_DIST_FORCE_STD] = B200_DIV(B200_MUL(B200_MUL(multiple, t2w), p[P_MASS]), 3.0f);
        }
        if(p[P_DR_TAU_RISE_MIN] == 0.0f || p[P_DR_TAU_RISE_MAX] == 0.0f || p[P_DR_TAU_FALL_MIN] == 0.0f || p[P_DR_TAU_FALL_MAX] == 0.0f) return false;
        {
            const float rising = rng_uniform(rng, p[P_DR_TAU_RISE_MIN], p[P_DR_TAU_RISE_MAX]);
            const float falling = rng_uniform(rng, p[P_DR_TAU_FALL_MIN], p[P_DR_TAU_FALL_MAX]);
            for(int r = 0; r < 4; r++){ p[P_TAU_RISE + r] = rising; p[P_TAU_FALL + r] = falling; }
        }
        return true;
    }
}

// hist_ptr: SoA rows of action_history for H > 1 (element (h,a) at hist_ptr[(4h+a)*n]); unused for H == 1
template <class Spec, class P>
__device__ __forceinline__ void set_history_from_rpm(EnvState<Spec>& st, const P& p, float* __restrict__ hist_ptr, size_t n){
    float v[4];
#pragma unroll
    for(int i = 0; i < 4; i++) v[i] = B200_SUB(B200_MUL(B200_DIV(B200_SUB(st.x[X_RPM + i], p[P_ACT_MIN]), B200_SUB(p[P_ACT_MAX], p[P_ACT_MIN])), 2.0f), 1.0f);
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) st.hist[i] = v[i];
    }
    else{
        for(int h = 0; h < Spec::H; h++){
#pragma unroll
            for(int i = 0; i < 4; i++) hist_ptr[(size_t)(4 * h + i) * n] = v[i];
        }
    }
    st.current_step = 0;
}

template <class Spec, class P>
__device__ __forceinline__ void initial_state(EnvState<Spec>& st, const P& p, float* __restrict__ hist_ptr, size_t n){
#pragma unroll
    for(int i = 0; i < X_DIM; i++) st.x[i] = 0.0f;
    st.x[X_ORI] = 1.0f;
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = 0.0f;
#pragma unroll
    for(int i = 0; i < 3; i++){ st.force[i] = 0.0f; st.torque[i] = 0.0f; }
#pragma unroll
    for(int i = 0; i < 4; i++) st.x[X_RPM + i] = B200_ADD(B200_MUL(p[P_HOVER], B200_SUB(p[P_ACT_MAX], p[P_ACT_MIN])), p[P_ACT_MIN]);
    set_history_from_rpm(st, p, hist_ptr, n);
    st.traj_type = 0;
    if constexpr(Spec::LANGEVIN){
#pragma unroll
        for(int i = 0; i < 12; i++) st.lang[i] = 0.0f;
    }
}

template <class Spec, class P>
__device__ __forceinline__ void sample_state(EnvState<Spec>& st, const P& p, uint64_t& rng, float* __restrict__ hist_ptr, size_t n){
#pragma unroll
    for(int i = 0; i < X_DIM; i++) st.x[i] = 0.0f;
    const bool guidance = rng_uniform(rng, 0.0f, 1.0f) < p[P_INIT_GUIDANCE];
    if(!guidance){
        const float mp = p[P_INIT_MAX_POS];
        for(int i = 0; i < 3; i++) st.x[X_POS + i] = rng_uniform(rng, -mp, mp);
    }
    if(p[P_INIT_MAX_ANGLE] > 0.0f && !guidance){
        const float u = rng_uniform(rng, 0.0f, 1.0f);
        const float v = rng_uniform(rng, 0.0f, 1.0f);
        const float phi = (float)(2.0 * (double)3.14159274101257324f * (double)u);
        const float cos_theta = (float)(1.0 - 2.0 * (double)v);
        const float sin_theta = (float)sqrt(1.0 - (double)B200_MUL(cos_theta, cos_theta));
        const float ax = B200_MUL(sin_theta, cosf(phi));
        const float ay = B200_MUL(sin_theta, sinf(phi));
        const float az = cos_theta;
        const float angle = rng_uniform(rng, 0.0f, 1.0f);   // the limit only gates (30_sample_initial_state.h:31,60)
        const float half = 0.5f * angle;
        const float sn = sinf(half);
        st.x[X_ORI] = cosf(half); st.x[X_ORI + 1] = B200_MUL(ax, sn); st.x[X_ORI + 2] = B200_MUL(ay, sn); st.x[X_ORI + 3] = B200_MUL(az, sn);
    }
    else{
        st.x[X_ORI] = 1.0f;
    }
    if(!guidance){
        const float mv = p[P_INIT_MAX_LINVEL], mw = p[P_INIT_MAX_ANGVEL];
        for(int i = 0; i < 3; i++) st.x[X_VEL + i] = rng_uniform(rng, -mv, mv);
        for(int i = 0; i < 3; i++) st.x[X_OMEGA + i] = rng_uniform(rng, -mw, mw);
    }
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = 0.0f;
    {
        const float fm = p[P_DIST_FORCE_MEAN], fs = p[P_DIST_FORCE_STD], tm = p[P_DIST_TORQUE_MEAN], ts = p[P_DIST_TORQUE_STD];
        for(int i = 0; i < 3; i++) st.force[i] = rng_normal(rng, fm, fs);
        st.torque[0] = rng_normal(rng, tm, ts);
        st.torque[1] = rng_normal(rng, tm, ts);
        st.torque[2] = rng_normal(rng, tm, B200_DIV(ts, 100.0f));
    }
    {
        float min_rpm, max_rpm;
        const float amin = p[P_ACT_MIN], amax = p[P_ACT_MAX];
        if(p[P_INIT_REL_RPM] != 0.0f){
            min_rpm = B200_ADD(B200_MUL(B200_DIV(B200_ADD(p[P_INIT_MIN_RPM], 1.0f), 2.0f), B200_SUB(amax, amin)), amin);
            max_rpm = B200_ADD(B200_MUL(B200_DIV(B200_ADD(p[P_INIT_MAX_RPM], 1.0f), 2.0f), B200_SUB(amax, amin)), amin);
        }
        else{
            min_rpm = p[P_INIT_MIN_RPM] < 0.0f ? amin : p[P_INIT_MIN_RPM];
            max_rpm = p[P_INIT_MAX_RPM] < 0.0f ? amax : p[P_INIT_MAX_RPM];
            if(max_rpm > amax) max_rpm = amax;
            if(min_rpm > max_rpm) min_rpm = max_rpm;
        }
        for(int i = 0; i < 4; i++) st.x[X_RPM + i] = rng_uniform(rng, min_rpm, max_rpm);
    }
    set_history_from_rpm(st, p, hist_ptr, n);
    st.traj_type = 0;
    if constexpr(Spec::LANGEVIN){
        const float threshold = rng_uniform(rng, 0.0f, 1.0f);
        float acc = 0.0f;
        int type = 0;
        for(int t = 0; t < 2; t++){
            acc = B200_ADD(acc, p[P_TRAJ_MIX0 + t]);
            if(threshold < acc){ type = t; break; }
        }
        st.traj_type = type;
#pragma unroll
        for(int i = 0; i < 12; i++) st.lang[i] = 0.0f;
    }
}

}  // namespace b200l2f
