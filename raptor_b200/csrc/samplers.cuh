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
// FAST twins for the in-kernel resets of the default-math kernels (k_collect_ts, k_off_policy_ts): MUFU reciprocal / cube root / square root / sine /
// cosine and the MUFU Box-Muller instead of IEEE division, libdevice cbrtf / sinf / cosf / logf and fp64 detours.  Every result stays within ~5e-7
// relative of the accurate form (the parity tests bound the re-sampled parameters at 2e-6); the integer RNG stream is untouched, so is the draw order.
template <bool FAST> __device__ __forceinline__ float div_t(float a, float b){ if constexpr(FAST) return a * rcp_approx(b); else return __fdiv_rn(a, b); }
template <bool FAST> __device__ __forceinline__ float cbrt_t(float x){   // x > 0 (masses)
    if constexpr(FAST){ float l; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x)); return ex2_approx(l * 0.333333343267440796f); }
    else return cbrtf(x);
}

// Register overlay over the (uniform) nominal row: the 47 entries domain randomisation rewrites live in registers, everything else is read
// from the row.  The sampler therefore issues no dependent global loads; the column in HBM is written once at the end (flush).  All indices
// reach get/set as compile-time constants (unrolled loops), so the selection chains fold away.
__host__ __device__ constexpr bool dr_overlay_index(int i){   // the entries sample_initial_parameters may rewrite
    return i == P_MASS || (i >= P_THRUST_COEF && i < P_THRUST_COEF + 12) || (i >= P_ROTOR_POS && i < P_ROTOR_POS + 12) ||
           (i >= P_J && i < P_J + 9 && (i - P_J) % 4 == 0) || (i >= P_JINV && i < P_JINV + 9 && (i - P_JINV) % 4 == 0) ||
           i == P_TERM_POS || i == P_INIT_MAX_POS || (i >= P_TORQUE_CONST && i < P_TORQUE_CONST + 4) || i == P_DIST_FORCE_MEAN || i == P_DIST_FORCE_STD ||
           (i >= P_TAU_RISE && i < P_TAU_RISE + 4) || (i >= P_TAU_FALL && i < P_TAU_FALL + 4);
}
struct ParamsOverlay {
    const float* __restrict__ row;
    float mass, coef[12], jd[3], jinvd[3], rpos[12], term_pos, init_max_pos, kq[4], dist_f_mean, dist_f_std, tau_rise[4], tau_fall[4];
    __device__ __forceinline__ void init(const float* __restrict__ r){
        row = r;
        mass = r[P_MASS]; term_pos = r[P_TERM_POS]; init_max_pos = r[P_INIT_MAX_POS]; dist_f_mean = r[P_DIST_FORCE_MEAN]; dist_f_std = r[P_DIST_FORCE_STD];
#pragma unroll
        for(int i = 0; i < 12; i++){ coef[i] = r[P_THRUST_COEF + i]; rpos[i] = r[P_ROTOR_POS + i]; }
#pragma unroll
        for(int i = 0; i < 3; i++){ jd[i] = r[P_J + 4 * i]; jinvd[i] = r[P_JINV + 4 * i]; }
#pragma unroll
        for(int i = 0; i < 4; i++){ kq[i] = r[P_TORQUE_CONST + i]; tau_rise[i] = r[P_TAU_RISE + i]; tau_fall[i] = r[P_TAU_FALL + i]; }
    }
    static __host__ __device__ constexpr bool diag(int i, int base){ return i >= base && i < base + 9 && (i - base) % 4 == 0; }
    __device__ __forceinline__ float operator[](int i) const {
        if(i == P_MASS) return mass;
        if(i >= P_THRUST_COEF && i < P_THRUST_COEF + 12) return coef[i - P_THRUST_COEF];
        if(i >= P_ROTOR_POS && i < P_ROTOR_POS + 12) return rpos[i - P_ROTOR_POS];
        if(diag(i, P_J)) return jd[(i - P_J) / 4];
        if(diag(i, P_JINV)) return jinvd[(i - P_JINV) / 4];
        if(i == P_TERM_POS) return term_pos;
        if(i == P_INIT_MAX_POS) return init_max_pos;
        if(i >= P_TORQUE_CONST && i < P_TORQUE_CONST + 4) return kq[i - P_TORQUE_CONST];
        if(i == P_DIST_FORCE_MEAN) return dist_f_mean;
        if(i == P_DIST_FORCE_STD) return dist_f_std;
        if(i >= P_TAU_RISE && i < P_TAU_RISE + 4) return tau_rise[i - P_TAU_RISE];
        if(i >= P_TAU_FALL && i < P_TAU_FALL + 4) return tau_fall[i - P_TAU_FALL];
        return row[i];
    }
    // the column in HBM: the row, then the overlaid entries (same thread, same addresses: program order).  FULL = false: the caller knows
    // that the column's other entries already equal the row (it was filled from the same row before), only the overlay is written.
    template <bool FULL = true>
    __device__ __forceinline__ void flush(const ParamsRW& g) const {
        if constexpr(FULL){
#pragma unroll 5
            for(int i = 0; i < PARAMS_DIM; i++) g[i] = row[i];
        }
        g[P_MASS] = mass; g[P_TERM_POS] = term_pos; g[P_INIT_MAX_POS] = init_max_pos; g[P_DIST_FORCE_MEAN] = dist_f_mean; g[P_DIST_FORCE_STD] = dist_f_std;
#pragma unroll
        for(int i = 0; i < 12; i++){ g[P_THRUST_COEF + i] = coef[i]; g[P_ROTOR_POS + i] = rpos[i]; }
#pragma unroll
        for(int i = 0; i < 3; i++){ g[P_J + 4 * i] = jd[i]; g[P_JINV + 4 * i] = jinvd[i]; }
#pragma unroll
        for(int i = 0; i < 4; i++){ g[P_TORQUE_CONST + i] = kq[i]; g[P_TAU_RISE + i] = tau_rise[i]; g[P_TAU_FALL + i] = tau_fall[i]; }
    }
};

// sample_initial_parameters on the overlay (p.init(row) done by the caller).  Returns false if the DR ranges violate the reference's
// assert_exit conditions (the caller records the error).  Arithmetic and draw order: 10_sample_initial_parameters.h:20-160.
template <bool DR, bool RNG_OOL = false, bool FAST = false>
__device__ __forceinline__ bool sample_parameters(ParamsOverlay& p, uint64_t& rng){
    if constexpr(!DR){ return true; }
    else{
        const float* __restrict__ row = p.row;
        float t2w_nominal;
        float gravity_norm = sqrt_t<FAST>(B200_ADD(B200_ADD(B200_MUL(row[P_GRAVITY], row[P_GRAVITY]), B200_MUL(row[P_GRAVITY + 1], row[P_GRAVITY + 1])), B200_MUL(row[P_GRAVITY + 2], row[P_GRAVITY + 2])));
        {
            const float max_action = row[P_ACT_MAX];
            float max_thrust_nominal = 0.0f;
#pragma unroll
            for(int r = 0; r < 4; r++){
                float v = B200_ADD(B200_ADD(p.coef[3 * r], B200_MUL(p.coef[3 * r + 1], max_action)), B200_MUL(B200_MUL(p.coef[3 * r + 2], max_action), max_action));
                max_thrust_nominal = B200_ADD(max_thrust_nominal, v);
            }
            t2w_nominal = div_t<FAST>(max_thrust_nominal, B200_MUL(p.mass, gravity_norm));
        }
        if(!(row[P_DR_T2W_MIN] < row[P_DR_T2W_MAX]) || !(row[P_DR_T2W_MIN] >= 1.5f)) return false;
        const float t2w = rng_uniform(rng, row[P_DR_T2W_MIN], row[P_DR_T2W_MAX]);
        const float factor_t2w = div_t<FAST>(t2w, t2w_nominal);
        if(!(row[P_DR_MASS_MIN] < row[P_DR_MASS_MAX])) return false;
        const float size_min = cbrt_t<FAST>(row[P_DR_MASS_MIN]);
        const float size_max = cbrt_t<FAST>(row[P_DR_MASS_MAX]);
        const float size_new = rng_uniform(rng, size_min, size_max);
        float mass_new = B200_MUL(B200_MUL(size_new, size_new), size_new);
        mass_new = clampf(mass_new, row[P_DR_MASS_MIN], row[P_DR_MASS_MAX]);
        const float scale_relative = cbrt_t<FAST>(div_t<FAST>(mass_new, p.mass));
        const float factor_mass = div_t<FAST>(mass_new, p.mass);
        p.mass = mass_new;
        const float factor_coef = B200_MUL(factor_t2w, factor_mass);
#pragma unroll
        for(int i = 0; i < 12; i++) p.coef[i] = B200_MUL(p.coef[i], factor_coef);
        float t2i_factor;
        {
            const float max_thrust = div_t<FAST>(B200_MUL(B200_MUL(t2w, p.mass), gravity_norm), 4.0f);
            const float first_rotor_distance = fabsf(p.rpos[0]);
            const float max_torque = FAST ? first_rotor_distance * 1.41421353816986084f * max_thrust : (float)((double)first_rotor_distance * 1.414213562373095 * (double)max_thrust);
            const float t2i_nominal = div_t<FAST>(max_torque, p.jd[0]);
            if(!(row[P_DR_T2I_MIN] < row[P_DR_T2I_MAX])) return false;
            const float t2i = rng_uniform(rng, row[P_DR_T2I_MIN], row[P_DR_T2I_MAX]);
            t2i_factor = div_t<FAST>(t2i, t2i_nominal);
        }
        if(row[P_DR_MASS_SIZE_DEV] == 0.0f) return false;
        float size_factor;
        {
            const float range = row[P_DR_MASS_SIZE_DEV];
            const float f = rng_normal_t<RNG_OOL, FAST>(rng, -range, range);
            size_factor = f < 0.0f ? div_t<FAST>(1.0f, B200_SUB(1.0f, f)) : B200_ADD(1.0f, f);
        }
        const float rotor_distance_factor = B200_MUL(scale_relative, size_factor);
        {
            const float inertia_factor = div_t<FAST>(t2i_factor, rotor_distance_factor);
#pragma unroll
            for(int a = 0; a < 3; a++){
                p.jd[a] = div_t<FAST>(p.jd[a], inertia_factor);
                p.jinvd[a] = B200_MUL(p.jinvd[a], inertia_factor);
            }
#pragma unroll
            for(int i = 0; i < 12; i++) p.rpos[i] = B200_MUL(p.rpos[i], rotor_distance_factor);
            float max_rotor_distance = 0.0f;
#pragma unroll
            for(int r = 0; r < 4; r++){
                const float x = p.rpos[3 * r], y = p.rpos[3 * r + 1], z = p.rpos[3 * r + 2];
                const float dd = sqrt_t<FAST>(B200_ADD(B200_ADD(B200_MUL(x, x), B200_MUL(y, y)), B200_MUL(z, z)));
                if(dd > max_rotor_distance) max_rotor_distance = dd;
            }
            p.term_pos = B200_MUL(max_rotor_distance, 20.0f);
            p.init_max_pos = B200_MUL(max_rotor_distance, 10.0f);
        }
        if(row[P_DR_KQ_MIN] == 0.0f || row[P_DR_KQ_MAX] == 0.0f) return false;
        {
            const float kq = rng_uniform(rng, row[P_DR_KQ_MIN], row[P_DR_KQ_MAX]);
#pragma unroll
            for(int r = 0; r < 4; r++) p.kq[r] = kq;
        }
        if(row[P_DR_DIST_FORCE_MAX] == 0.0f) return false;
        {
            float surplus = FAST ? t2w - 1.0f : (float)((double)t2w - 1.0);
            if(surplus < 0.0f) surplus = 0.0f;
            const float multiple = rng_uniform(rng, 0.0f, B200_MUL(surplus, row[P_DR_DIST_FORCE_MAX]));
            p.dist_f_mean = 0.0f;
            p.dist_f_std = div_t<FAST>(B200_MUL(B200_MUL(multiple, t2w), p.mass), 3.0f);
        }
        if(row[P_DR_TAU_RISE_MIN] == 0.0f || row[P_DR_TAU_RISE_MAX] == 0.0f || row[P_DR_TAU_FALL_MIN] == 0.0f || row[P_DR_TAU_FALL_MAX] == 0.0f) return false;
        {
            const float rising = rng_uniform(rng, row[P_DR_TAU_RISE_MIN], row[P_DR_TAU_RISE_MAX]);
            const float falling = rng_uniform(rng, row[P_DR_TAU_FALL_MIN], row[P_DR_TAU_FALL_MAX]);
#pragma unroll
            for(int r = 0; r < 4; r++){ p.tau_rise[r] = rising; p.tau_fall[r] = falling; }
        }
        return true;
    }
}
// row -> sampled column in HBM.  On a range error the column holds the state at the point of the error (the caller raises the error).
template <bool DR>
__device__ __forceinline__ bool sample_parameters(const float* __restrict__ env_p, const ParamsRW& out, uint64_t& rng){
    ParamsOverlay o;
    o.init(env_p);
    const bool ok = sample_parameters<DR>(o, rng);
    o.flush(out);
    return ok;
}

// hist_ptr: SoA rows of action_history for H > 1 (element (h,a) at hist_ptr[(4h+a)*n]); unused for H == 1
template <class Spec, class P, bool FAST = false>
__device__ __forceinline__ void set_history_from_rpm(EnvState<Spec>& st, const P& p, float* __restrict__ hist_ptr, size_t n){
    float v[4];
#pragma unroll
    for(int i = 0; i < 4; i++) v[i] = B200_SUB(B200_MUL(div_t<FAST>(B200_SUB(st.x[X_RPM + i], p[P_ACT_MIN]), B200_SUB(p[P_ACT_MAX], p[P_ACT_MIN])), 2.0f), 1.0f);
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

// KEEP_DEAD_TARGET: a reset that draws the POSITION trajectory leaves the previous episode's Langevin target in the state, exactly as the
// reference does (30_sample_initial_state.h:217-219: `case POSITION: break;`); the values are never read while the type is POSITION and are
// zeroed when the type becomes LANGEVIN again.  false: the state is built from scratch (the vector-API call on a fresh State).
template <class Spec, class P, bool KEEP_DEAD_TARGET = false, bool FAST = false>
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
        float ax, ay, az, angle;
        if constexpr(FAST){                                  // same quantities on the MUFU pipe: phi reduced to (-pi, pi], sin(theta) = sqrt((1 - c)(1 + c)) without the fp64 detour
            const float phi = 6.28318548202514648f * (u > 0.5f ? u - 1.0f : u);
            const float cos_theta = 1.0f - 2.0f * v;
            const float sin_theta = sqrt_approx((1.0f - cos_theta) * (1.0f + cos_theta));
            ax = sin_theta * __cosf(phi); ay = sin_theta * __sinf(phi); az = cos_theta;
        }
        else{
            const float phi = (float)(2.0 * (double)3.14159274101257324f * (double)u);
            const float cos_theta = (float)(1.0 - 2.0 * (double)v);
            const float sin_theta = (float)sqrt(1.0 - (double)B200_MUL(cos_theta, cos_theta));
            ax = B200_MUL(sin_theta, cosf(phi));
            ay = B200_MUL(sin_theta, sinf(phi));
            az = cos_theta;
        }
        angle = rng_uniform(rng, 0.0f, 1.0f);               // the limit only gates (30_sample_initial_state.h:31,60)
        const float half = 0.5f * angle;
        const float sn = FAST ? __sinf(half) : sinf(half);
        st.x[X_ORI] = FAST ? __cosf(half) : cosf(half); st.x[X_ORI + 1] = B200_MUL(ax, sn); st.x[X_ORI + 2] = B200_MUL(ay, sn); st.x[X_ORI + 3] = B200_MUL(az, sn);
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
        for(int i = 0; i < 3; i++) st.force[i] = rng_normal_t<Spec::RNG_OOL, FAST>(rng, fm, fs);
        st.torque[0] = rng_normal_t<Spec::RNG_OOL, FAST>(rng, tm, ts);
        st.torque[1] = rng_normal_t<Spec::RNG_OOL, FAST>(rng, tm, ts);
        st.torque[2] = rng_normal_t<Spec::RNG_OOL, FAST>(rng, tm, div_t<FAST>(ts, 100.0f));
    }
    {
        float min_rpm, max_rpm;
        const float amin = p[P_ACT_MIN], amax = p[P_ACT_MAX];
        if(p[P_INIT_REL_RPM] != 0.0f){
            min_rpm = B200_ADD(B200_MUL(B200_MUL(B200_ADD(p[P_INIT_MIN_RPM], 1.0f), 0.5f), B200_SUB(amax, amin)), amin);   // / 2.0f == * 0.5f exactly
            max_rpm = B200_ADD(B200_MUL(B200_MUL(B200_ADD(p[P_INIT_MAX_RPM], 1.0f), 0.5f), B200_SUB(amax, amin)), amin);
        }
        else{
            min_rpm = p[P_INIT_MIN_RPM] < 0.0f ? amin : p[P_INIT_MIN_RPM];
            max_rpm = p[P_INIT_MAX_RPM] < 0.0f ? amax : p[P_INIT_MAX_RPM];
            if(max_rpm > amax) max_rpm = amax;
            if(min_rpm > max_rpm) min_rpm = max_rpm;
        }
        for(int i = 0; i < 4; i++) st.x[X_RPM + i] = rng_uniform(rng, min_rpm, max_rpm);
    }
    set_history_from_rpm<Spec, P, FAST>(st, p, hist_ptr, n);
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
        if(!KEEP_DEAD_TARGET || type == 1){
#pragma unroll
            for(int i = 0; i < 12; i++) st.lang[i] = 0.0f;
        }
    }
}

}  // namespace b200l2f
