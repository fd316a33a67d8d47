// raptor_b200/csrc/rollout_x2.cuh -- k_rollout_raptor_x2: the fused persistent rollout with TWO environments per thread on the packed
// fp32 pipe of sm_100 (FFMA2 / FADD2 / FMUL2: one instruction = the same operation on two fp32 lanes).
//
// k_rollout_raptor_ts (rollout_tc.cuh) is bound by instruction issue: 1416 warp instructions per 32 environment steps, of which ~530 are
// scalar fp32 arithmetic of the environment (RK4 dynamics, observation, reward).  Here a thread owns environments A and B = A + 128 and
// carries every integrated quantity as a float2 (lane x = A, lane y = B), so that the environment's arithmetic issues once for both:
//   * one CTA = 128 threads = 256 environments = TWO M = 128 MMA tiles (A rows in TMEM columns [0, 128), B rows in [128, 256)); the
//     actor GEMMs, the TMEM-resident GRU state and the 3xTF32 split are those of the TS kernel, issued for both tiles behind ONE barrier /
//     commit / wait per GEMM stage (half the synchronisation per environment step);
//   * per-environment dynamics constants are staged pair-interleaved, sm[tid][i][e], so one LDS.128 yields entries i, i + 1 of both
//     environments as two packed operands; loop invariants (1/m, g + F_d/m, J^-1 tau_d, dt/2, dt/3, dt/6, sqrt(dt), the motor lag as
//     m d + h |d|) are compiled into the block once per (tile pair, time chunk);
//   * what does not pack -- MUFU, the 64-bit xorshift stream, min/max, the TMEM loads / stores -- runs per environment; the GRU gate
//     epilogue and dense 2 stay packed over two hidden units of ONE environment as in the TS kernel (TMEM loads deliver consecutive units
//     in consecutive registers);
//   * 2 CTAs (8 warps, 512 environments) per SM: 255 registers, 102 KB of shared memory, 256 TMEM columns per CTA.
// Specialised for the configuration the hot path is benchmarked on (foundation-policy spec: H = 1, Langevin targets, OBS 22; default math;
// uniform MDP constants; axial vehicles; no observation / action noise); everything else takes k_rollout_raptor_ts.
// Arithmetic follows the same reference functions as env.cuh / rollout_tc.cuh (cited there); association differs from the one-environment
// kernels only inside fused multiply-adds, so the two agree to fp32 rounding, not bit for bit.
#pragma once
#include <type_traits>
#include "rollout_tc.cuh"

namespace b200l2f {

// ---- compiled per-environment block, pair-interleaved: entry i of environment e at sm[2 i + e] of the thread's row ------------------------
enum DynX2 : int {
    X2_COEF = 0,        // [12] thrust curve, rotor r at 3 r .. 3 r + 2
    X2_AT = 12,         // [3][4] torque per unit rotor thrust
    X2_TAU_M = 24,      // [4] (1/tau_rise + 1/tau_fall) / 2
    X2_TAU_H = 28,      // [4] (1/tau_rise - 1/tau_fall) / 2        d(rpm)/dt = m d + h |d|, d = setpoint - rpm
    X2_GA = 32,         // [3] gravity + F_d / m                    | 35: 1 / m
    X2_INV_MASS = 35,
    X2_JD = 36,         // [3] diag(J)                              | 39: action min
    X2_AMIN = 39,
    X2_JID = 40,        // [3] diag(J^-1)                           | 43: action max
    X2_AMAX = 43,
    X2_TA = 44,         // [3] J^-1 tau_d                           | 47: (max - min) / 2
    X2_HALF_RANGE = 47,
    X2_DT = 48, X2_DT2 = 49, X2_DT3 = 50, X2_DT6 = 51,
    X2_SQRT_DT = 52, X2_TERM_POS = 53, X2_SP_OFFSET = 54,           // 54: min + (max - min) / 2
    X2_DIM = 58         // row stride 2 x 58 = 116 words = 20 mod 32: the eight threads of a quarter-warp LDS.128 hit disjoint banks
};
static_assert((2 * X2_DIM) % 4 == 0 && ((2 * X2_DIM) % 32 == 4 || (2 * X2_DIM) % 32 == 12 || (2 * X2_DIM) % 32 == 20 || (2 * X2_DIM) % 32 == 28), "conflict-free LDS.128 row stride");

struct X2Smem {
    static constexpr int B = 0;                                        // weight image (TMA destination)
    static constexpr int DYN = B + TcImage::BYTES;
    static constexpr int LANG = DYN + 2 * X2_DIM * BLOCK * 4;          // Langevin target state, per environment 3 float4: [(e * 3 + k) * BLOCK + tid]
    static constexpr int LAST = LANG + 6 * BLOCK * 16;                 // last action, per environment one float4: [e * BLOCK + tid]
    static constexpr int BAR = LAST + 2 * BLOCK * 16;
    static constexpr int TOTAL = BAR + 32;
    static constexpr int ENVS = 2 * BLOCK;                             // environments per CTA (tile pair)
};
static_assert(X2Smem::DYN % 16 == 0 && X2Smem::LANG % 16 == 0, "float4 alignment");

struct BlockX2 {
    const float* sm;   // this thread's row
    __device__ __forceinline__ float4 q(int i) const { return *reinterpret_cast<const float4*>(sm + 2 * i); }   // i even: entries i, i + 1 of both environments
    __device__ __forceinline__ F2 p(int i) const { return *reinterpret_cast<const F2*>(sm + 2 * i); }
};

// multirotor dynamics of two axial vehicles (same physics as dynamics_compiled<AXIAL = true>; 60_dynamics.h:18-72, :87-111)
__device__ __forceinline__ void dynamics_x2(const BlockX2& b, const F2* __restrict__ x, const F2* __restrict__ sp, F2* __restrict__ dx){
    using namespace p2;
    F2 tm[4];
    {
        const float4 c0 = b.q(X2_COEF), c1 = b.q(X2_COEF + 2), c2 = b.q(X2_COEF + 4), c3 = b.q(X2_COEF + 6), c4 = b.q(X2_COEF + 8), c5 = b.q(X2_COEF + 10);
        const F2 cf[12] = {lo2(c0), hi2(c0), lo2(c1), hi2(c1), lo2(c2), hi2(c2), lo2(c3), hi2(c3), lo2(c4), hi2(c4), lo2(c5), hi2(c5)};
#pragma unroll
        for(int r = 0; r < 4; r++){
            const F2 rpm = x[X_RPM + r];
            tm[r] = fma(fma(cf[3 * r + 2], rpm, cf[3 * r + 1]), rpm, cf[3 * r]);
        }
    }
    const F2 T = add(add(add(tm[0], tm[1]), tm[2]), tm[3]);
    F2 torque[3];
#pragma unroll
    for(int i = 0; i < 3; i++){
        const float4 u = b.q(X2_AT + 4 * i), v = b.q(X2_AT + 4 * i + 2);
        torque[i] = fma(hi2(v), tm[3], fma(lo2(v), tm[2], fma(hi2(u), tm[1], mul(lo2(u), tm[0]))));
    }
#pragma unroll
    for(int i = 0; i < 3; i++) dx[X_POS + i] = x[X_VEL + i];
    const F2 q0 = x[X_ORI], q1 = x[X_ORI + 1], q2 = x[X_ORI + 2], q3 = x[X_ORI + 3];
    const F2 w0 = x[X_OMEGA], w1 = x[X_OMEGA + 1], w2 = x[X_OMEGA + 2];
    {
        const F2 half = bc(0.5f);
        const F2 h0 = mul(w0, half), h1 = mul(w1, half), h2 = mul(w2, half);   // exact scaling: same values as (...) * 0.5
        dx[X_ORI + 0] = fnma(q3, h2, fnma(q2, h1, mul(neg(q1), h0)));
        dx[X_ORI + 1] = fnma(q3, h1, fma(q2, h2, mul(q0, h0)));
        dx[X_ORI + 2] = fnma(q1, h2, fma(q3, h0, mul(q0, h1)));
        dx[X_ORI + 3] = fnma(q2, h0, fma(q1, h1, mul(q0, h2)));
    }
    {   // rotate (0, 0, T) by q, / m, + g + F_d / m
        const F2 T2 = add(T, T);
        const F2 v0 = mul(q2, T2), v1 = mul(neg(q1), T2);
        const F2 o0 = fma(v0, q0, mul(neg(q3), v1));
        const F2 o1 = fma(v1, q0, mul(q3, v0));
        const F2 o2 = add(fnma(q2, v0, mul(q1, v1)), T);
        const float4 g0 = b.q(X2_GA), g1 = b.q(X2_GA + 2);
        const F2 im = hi2(g1);
        dx[X_VEL + 0] = fma(o0, im, lo2(g0));
        dx[X_VEL + 1] = fma(o1, im, hi2(g0));
        dx[X_VEL + 2] = fma(o2, im, lo2(g1));
    }
    {   // J^-1 (tau - w x J w) + J^-1 tau_d
        const float4 j0 = b.q(X2_JD), j1 = b.q(X2_JD + 2), i0 = b.q(X2_JID), i1 = b.q(X2_JID + 2), t0 = b.q(X2_TA), t1 = b.q(X2_TA + 2);
        const F2 v0 = mul(lo2(j0), w0), v1 = mul(hi2(j0), w1), v2 = mul(lo2(j1), w2);
        const F2 e0 = fnma(w1, v2, fma(w2, v1, torque[0]));
        const F2 e1 = fnma(w2, v0, fma(w0, v2, torque[1]));
        const F2 e2 = fnma(w0, v1, fma(w1, v0, torque[2]));
        dx[X_OMEGA + 0] = fma(lo2(i0), e0, lo2(t0));
        dx[X_OMEGA + 1] = fma(hi2(i0), e1, hi2(t0));
        dx[X_OMEGA + 2] = fma(lo2(i1), e2, lo2(t1));
    }
    {   // first-order motor lag, rising / falling time constants: (setpoint - rpm) / tau(sign) = m d + h |d|
        const float4 m0 = b.q(X2_TAU_M), m1 = b.q(X2_TAU_M + 2), h0 = b.q(X2_TAU_H), h1 = b.q(X2_TAU_H + 2);
        const F2 m[4] = {lo2(m0), hi2(m0), lo2(m1), hi2(m1)}, h[4] = {lo2(h0), hi2(h0), lo2(h1), hi2(h1)};
#pragma unroll
        for(int r = 0; r < 4; r++){
            const F2 d = sub(sp[r], x[X_RPM + r]);
            dx[X_RPM + r] = fma(h[r], abs2(d), mul(m[r], d));
        }
    }
}

// acos(1 - u) for u in [0, 1]: sqrt(2 u) Q(u), Q a degree-7 minimax fit (|error| < 2.5e-7, fp32 Horner); replaces libdevice acosf in the
// orientation cost 2 acos(1 - |q_z|) (squared/operations_generic.h:24)
__device__ __forceinline__ F2 acos_1m_x2(F2 u){
    using namespace p2;
    F2 q = bc(8.746944950e-04f);
    q = fma(q, u, bc(-1.469472889e-03f)); q = fma(q, u, bc(2.444653539e-03f)); q = fma(q, u, bc(9.892442031e-04f)); q = fma(q, u, bc(5.829527043e-03f));
    q = fma(q, u, bc(1.871711761e-02f)); q = fma(q, u, bc(8.333496749e-02f)); q = fma(q, u, bc(1.0f));
    const F2 u2 = add(u, u);
    return mul(mk(sqrt_approx(u2.x), sqrt_approx(u2.y)), q);
}

template <class Spec>
__global__ void __launch_bounds__(BLOCK, 2) k_rollout_raptor_x2(const __grid_constant__ RolloutArgs a, const float* __restrict__ tc_image){
    static_assert(Spec::H == 1 && Spec::LANGEVIN && Spec::OBS_LAYOUT == OBS_RAPTOR, "two-environment kernel: foundation-policy specification only");
    using namespace p2;
    constexpr int HD = 16;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* sm_b = reinterpret_cast<float*>(smraw + X2Smem::B);
    float* sm_row = reinterpret_cast<float*>(smraw + X2Smem::DYN) + threadIdx.x * (2 * X2_DIM);
    float4* sm_lang = reinterpret_cast<float4*>(smraw + X2Smem::LANG) + threadIdx.x;
    float4* sm_last = reinterpret_cast<float4*>(smraw + X2Smem::LAST) + threadIdx.x;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smraw + X2Smem::BAR);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x;
    const int warp = tc::uniform_warp_index();   // TMEM addresses stay in uniform registers
    if(tid == 0){
        tc::mbar_init(bar_tma, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
    }
    if(warp == 0) tc::tmem_alloc<256>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if(tid == 0){
        tc::mbar_expect_tx(bar_tma, TcImage::BYTES);
        tc::tma_load_1d(sm_b, tc_image, TcImage::BYTES, bar_tma);
    }
    const size_t n = (size_t)a.n;
    const bool no_auto_reset = a.no_auto_reset != 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);    // this thread's TMEM lane, column 0 of tile A (tile B: + 128)
    __shared__ int s_item;
    constexpr uint32_t TILE = 128;   // TMEM columns per tile; plan inside a tile as in k_rollout_raptor_ts
    constexpr uint32_t C_H_HI = 0, C_H_LO = 16, C_OBS_HI = 32, C_OBS_LO = 56, C_D1 = 80, C_X1_HI = 32, C_X1_LO = 48, C_D2 = 64;
    auto put8 = [&](uint32_t col_hi, uint32_t col_lo, const float* v){   // columns already include the tile offset
        float hi[8], lo[8];
#pragma unroll
        for(int i = 0; i < 8; i++) tc::split_tf32(v[i], hi[i], lo[i]);
        tc::tmem_st8(lane_addr + col_hi, hi);
        tc::tmem_st8(lane_addr + col_lo, lo);
    };
    const uint32_t b_s = tc::smem_u32(sm_b);
    constexpr uint32_t SBO = 128;
    constexpr uint32_t IDESC16 = tc::make_idesc_tf32(128, 16), IDESC64 = tc::make_idesc_tf32(128, 64);
    uint32_t phase = 0;
    const uint64_t desc_n16 = tc::make_smem_desc(b_s, 16 * 16, SBO), desc_n64 = tc::make_smem_desc(b_s, 64 * 16, SBO);
    auto issue_gemm = [&](uint32_t dcol, uint32_t a_hi, uint32_t a_lo, int ksteps, int b_hi_off, int b_lo_off, int b_pair0, uint32_t N, uint32_t idesc, uint32_t acc){
        const uint64_t base = N == 16 ? desc_n16 : desc_n64;
#pragma unroll
        for(int s = 0; s < ksteps; s++){
            const uint64_t bhi = tc::smem_desc_advance(base, b_hi_off * 4 + (b_pair0 + s) * 2 * N * 16);
            const uint64_t blo = tc::smem_desc_advance(base, b_lo_off * 4 + (b_pair0 + s) * 2 * N * 16);
            tc::mma_tf32_ts(tmem_base + dcol, tmem_base + a_hi + 8 * s, bhi, idesc, acc); acc = 1;
            tc::mma_tf32_ts(tmem_base + dcol, tmem_base + a_hi + 8 * s, blo, idesc, 1);
            tc::mma_tf32_ts(tmem_base + dcol, tmem_base + a_lo + 8 * s, bhi, idesc, 1);
        }
    };
    const BlockX2 blk{sm_row};
    tc::mbar_wait(bar_tma, 0);
    __syncthreads();

    const int n_pairs = (a.n + X2Smem::ENVS - 1) / X2Smem::ENVS;
    const int n_chunks = a.n_chunks;
    const int total_items = n_pairs * n_chunks;
    // uniform MDP constants (launch constant bank)
    const float* R0 = a.row0;
    for(;;){
    if(tid == 0) s_item = atomicAdd(a.sched, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if(item >= total_items) break;
    const int pair = item % n_pairs, chunk = item / n_pairs;
    if(chunk > 0){
        if(tid == 0){
            const int* prog = a.sched + 1 + pair;
            int v;
            do{ asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog) : "memory"); if(v < chunk) __nanosleep(64); } while(v < chunk);
        }
        __syncthreads();
        __threadfence();
    }
    const int t_begin = chunk * a.chunk_steps;
    const int t_end = min(a.T, t_begin + a.chunk_steps);
    const int e0 = pair * X2Smem::ENVS + tid, e1 = e0 + BLOCK;
    const bool act0 = e0 < a.n, act1 = e1 < a.n;                 // inactive environments shadow environment 0 and never store
    const size_t env0 = act0 ? (size_t)e0 : 0, env1 = act1 ? (size_t)e1 : 0;
    auto env_of = [&](auto E) -> size_t { return decltype(E)::value == 0 ? env0 : env1; };
    auto active_of = [&](auto E) -> bool { return decltype(E)::value == 0 ? act0 : act1; };
    using E0 = std::integral_constant<int, 0>; using E1 = std::integral_constant<int, 1>;

    // ---- stage the compiled block of both environments
    auto stage = [&](auto E){
        constexpr int e = decltype(E)::value;
        const float* g = a.params + env_of(E);
        const float* s = a.state + env_of(E);
        auto P = [&](int i){ return __ldg(g + (size_t)i * n); };
        auto put = [&](int i, float v){ sm_row[2 * i + e] = v; };
#pragma unroll
        for(int i = 0; i < 12; i++) put(X2_COEF + i, P(P_THRUST_COEF + i));
#pragma unroll
        for(int r = 0; r < 4; r++){
            const float dx = P(P_THRUST_DIR + 3 * r), dy = P(P_THRUST_DIR + 3 * r + 1), dz = P(P_THRUST_DIR + 3 * r + 2);
            const float px = P(P_ROTOR_POS + 3 * r), py = P(P_ROTOR_POS + 3 * r + 1), pz = P(P_ROTOR_POS + 3 * r + 2);
            const float kq = P(P_TORQUE_CONST + r);
            put(X2_AT + 0 * 4 + r, P(P_TORQUE_DIR + 3 * r + 0) * kq + (py * dz - pz * dy));   // 60_dynamics.h:38-39
            put(X2_AT + 1 * 4 + r, P(P_TORQUE_DIR + 3 * r + 1) * kq + (pz * dx - px * dz));
            put(X2_AT + 2 * 4 + r, P(P_TORQUE_DIR + 3 * r + 2) * kq + (px * dy - py * dx));
            const float ir = 1.0f / P(P_TAU_RISE + r), ifl = 1.0f / P(P_TAU_FALL + r);
            put(X2_TAU_M + r, 0.5f * (ir + ifl)); put(X2_TAU_H + r, 0.5f * (ir - ifl));
        }
        const float mass = P(P_MASS);
        put(X2_INV_MASS, 1.0f / mass);
        float td[3];
#pragma unroll
        for(int i = 0; i < 3; i++){
            put(X2_GA + i, P(P_GRAVITY + i) + __ldcg(s + (size_t)(S_FORCE + i) * n) / mass);
            td[i] = __ldcg(s + (size_t)(S_TORQUE + i) * n);
            put(X2_JD + i, P(P_J + 4 * i)); put(X2_JID + i, P(P_JINV + 4 * i));
        }
#pragma unroll
        for(int i = 0; i < 3; i++) put(X2_TA + i, P(P_JINV + 3 * i) * td[0] + P(P_JINV + 3 * i + 1) * td[1] + P(P_JINV + 3 * i + 2) * td[2]);
        const float amin = P(P_ACT_MIN), amax = P(P_ACT_MAX), hr = (amax - amin) / 2.0f, dt = P(P_DT);
        put(X2_AMIN, amin); put(X2_AMAX, amax); put(X2_HALF_RANGE, hr); put(X2_SP_OFFSET, amin + hr);
        put(X2_DT, dt); put(X2_DT2, dt / 2.0f); put(X2_DT3, dt / 3.0f); put(X2_DT6, dt / 6.0f); put(X2_SQRT_DT, sqrtf(dt));
        put(X2_TERM_POS, P(P_TERM_POS));
    };
    stage(E0{}); stage(E1{});

    // ---- state of both environments
    const float* s0 = a.state + env0; const float* s1 = a.state + env1;
    auto ld2 = [&](int row){ return mk(__ldcg(s0 + (size_t)row * n), __ldcg(s1 + (size_t)row * n)); };
    F2 x[X_DIM], hist[4];
#pragma unroll
    for(int i = 0; i < 13; i++) x[i] = ld2(i);
#pragma unroll
    for(int i = 0; i < 4; i++){ x[X_RPM + i] = ld2(S_RPM + i); hist[i] = ld2(S_HIST + i); }
    auto park = [&](auto E){   // last action and Langevin state of environment E -> shared memory
        constexpr int e = decltype(E)::value;
        const float* s = e == 0 ? s0 : s1;
        sm_last[e * BLOCK] = make_float4(__ldcg(s + (size_t)(S_LAST_ACTION + 0) * n), __ldcg(s + (size_t)(S_LAST_ACTION + 1) * n), __ldcg(s + (size_t)(S_LAST_ACTION + 2) * n), __ldcg(s + (size_t)(S_LAST_ACTION + 3) * n));
#pragma unroll
        for(int k = 0; k < 3; k++){
            const int r = s_langevin(1) + 4 * k;
            sm_lang[(e * 3 + k) * BLOCK] = make_float4(__ldcg(s + (size_t)r * n), __ldcg(s + (size_t)(r + 1) * n), __ldcg(s + (size_t)(r + 2) * n), __ldcg(s + (size_t)(r + 3) * n));
        }
    };
    park(E0{}); park(E1{});
    const bool lang0 = (int)__ldcg(s0 + (size_t)s_traj_type(1) * n) == 1, lang1 = (int)__ldcg(s1 + (size_t)s_traj_type(1) * n) == 1;
    uint64_t rng0 = __ldcg(a.rng + env0), rng1 = __ldcg(a.rng + env1);
    int gs0 = __ldcg(a.gru_step + env0), gs1 = __ldcg(a.gru_step + env1);
    F2 ret = mk(0.0f, 0.0f); int eplen0 = 0, eplen1 = 0; bool done0 = false, done1 = false;
    if(chunk > 0){
        ret = mk(__ldcg(a.acc_ret + env0), __ldcg(a.acc_ret + env1));
        const int v0 = __ldcg(a.acc_len + env0), v1 = __ldcg(a.acc_len + env1);
        eplen0 = v0 >> 1; done0 = (v0 & 1) != 0; eplen1 = v1 >> 1; done1 = (v1 & 1) != 0;
    }
    {
        float h[HD];
#pragma unroll
        for(int j = 0; j < HD; j++) h[j] = __ldcg(a.hidden + (size_t)j * n + env0);
        put8(C_H_HI, C_H_LO, h); put8(C_H_HI + 8, C_H_LO + 8, h + 8);
#pragma unroll
        for(int j = 0; j < HD; j++) h[j] = __ldcg(a.hidden + (size_t)j * n + env1);
        put8(TILE + C_H_HI, TILE + C_H_LO, h); put8(TILE + C_H_HI + 8, TILE + C_H_LO + 8, h + 8);
    }
    tc::tmem_st_wait();
    // slow path: one state row of environment E in the reference layout (recording only)
    auto write_row = [&](auto E, float* __restrict__ row){
        constexpr int e = decltype(E)::value;
        const float* s = e == 0 ? s0 : s1;
#pragma unroll
        for(int i = 0; i < 13; i++) row[i] = lane<e>(x[i]);
        const float4 la = sm_last[e * BLOCK];
        row[S_LAST_ACTION] = la.x; row[S_LAST_ACTION + 1] = la.y; row[S_LAST_ACTION + 2] = la.z; row[S_LAST_ACTION + 3] = la.w;
#pragma unroll
        for(int i = 0; i < 3; i++){
            row[S_ANGVEL_HIST + i] = lane<e>(x[X_OMEGA + i]);
            row[S_FORCE + i] = __ldcg(s + (size_t)(S_FORCE + i) * n); row[S_TORQUE + i] = __ldcg(s + (size_t)(S_TORQUE + i) * n);
        }
#pragma unroll
        for(int i = 0; i < 4; i++){ row[S_RPM + i] = lane<e>(x[X_RPM + i]); row[S_HIST + i] = lane<e>(hist[i]); }
        row[S_CURRENT_STEP] = 0.0f;
        row[s_traj_type(1)] = (e == 0 ? lang0 : lang1) ? 1.0f : __ldcg(s + (size_t)s_traj_type(1) * n);
#pragma unroll
        for(int k = 0; k < 3; k++){
            const float4 v = sm_lang[(e * 3 + k) * BLOCK];
            float* d = row + s_langevin(1) + 4 * k;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    };

    for(int t = t_begin; t < t_end; t++){
        if(a.out_states && (t % a.state_stride) == 0){
            float* base = a.out_states + (size_t)(t / a.state_stride) * n * Spec::STATE_DIM;
            if(act0) write_row(E0{}, base + env0 * Spec::STATE_DIM);
            if(act1) write_row(E1{}, base + env1 * Spec::STATE_DIM);
        }
        // ---- observe (40_observe.h: position, rotation matrix, linear velocity, angular velocity, most recent action) + bias column
        {
            F2 o[24];
#pragma unroll
            for(int i = 0; i < 3; i++){ o[i] = x[X_POS + i]; o[12 + i] = x[X_VEL + i]; o[15 + i] = x[X_OMEGA + i]; }
            const F2 q0 = x[X_ORI], q1 = x[X_ORI + 1], q2 = x[X_ORI + 2], q3 = x[X_ORI + 3];
            const F2 d0 = add(q0, q0), d1 = add(q1, q1), d2 = add(q2, q2), d3 = add(q3, q3);
            const F2 one = bc(1.0f);
            const F2 p12 = mul(d1, q2), p13 = mul(d1, q3), p23 = mul(d2, q3);
            o[3]  = fnma(d3, q3, fnma(d2, q2, one));
            o[4]  = fnma(d0, q3, p12);
            o[5]  = fma(d0, q2, p13);
            o[6]  = fma(d0, q3, p12);
            o[7]  = fnma(d3, q3, fnma(d1, q1, one));
            o[8]  = fnma(d0, q1, p23);
            o[9]  = fnma(d0, q2, p13);
            o[10] = fma(d0, q1, p23);
            o[11] = fnma(d2, q2, fnma(d1, q1, one));
#pragma unroll
            for(int i = 0; i < 4; i++) o[18 + i] = hist[i];
            o[22] = one; o[23] = bc(0.0f);
            auto emit = [&](auto E){
                constexpr int e = decltype(E)::value;
                float v[24];
#pragma unroll
                for(int i = 0; i < 24; i++) v[i] = lane<e>(o[i]);
                if(a.out_obs && active_of(E)){
                    float* row = a.out_obs + ((size_t)t * n + env_of(E)) * 22;
#pragma unroll
                    for(int i = 0; i < 22; i++) row[i] = v[i];
                }
                put8(e * TILE + C_OBS_HI, e * TILE + C_OBS_LO, v); put8(e * TILE + C_OBS_HI + 8, e * TILE + C_OBS_LO + 8, v + 8); put8(e * TILE + C_OBS_HI + 16, e * TILE + C_OBS_LO + 16, v + 16);
            };
            emit(E0{}); emit(E1{});
        }
        // ---- G1: dense 1 of both tiles
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        if(warp == 0 && tc::elect_one()){
            tc::tc_fence_after();
            issue_gemm(C_D1, C_OBS_HI, C_OBS_LO, 3, TcImage::B1_HI, TcImage::B1_LO, 0, 16, IDESC16, 0);
            issue_gemm(TILE + C_D1, TILE + C_OBS_HI, TILE + C_OBS_LO, 3, TcImage::B1_HI, TcImage::B1_LO, 0, 16, IDESC16, 0);
            tc::mma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, phase); phase ^= 1;
        tc::tc_fence_after();
        {
            float xa[HD], xb[HD];
            tc::tmem_ld16(lane_addr + C_D1, xa);
            tc::tmem_ld16(lane_addr + TILE + C_D1, xb);
            tc::tmem_ld_wait();
#pragma unroll
            for(int j = 0; j < HD; j++){ xa[j] = xa[j] + fabsf(xa[j]); xb[j] = xb[j] + fabsf(xb[j]); }   // 2 ReLU: the scaled-gate image carries the 0.5 in its W_ih columns (build_tc_image_host)
            if(!no_auto_reset){   // reset_truncate (gru/operations_generic.h:76-86)
                if(gs0 >= a.seq_len){ put8(C_H_HI, C_H_LO, sm_b + TcImage::H0); put8(C_H_HI + 8, C_H_LO + 8, sm_b + TcImage::H0 + 8); gs0 = 0; }
                if(gs1 >= a.seq_len){ put8(TILE + C_H_HI, TILE + C_H_LO, sm_b + TcImage::H0); put8(TILE + C_H_HI + 8, TILE + C_H_LO + 8, sm_b + TcImage::H0 + 8); gs1 = 0; }
            }
            put8(C_X1_HI, C_X1_LO, xa); put8(C_X1_HI + 8, C_X1_LO + 8, xa + 8);
            put8(TILE + C_X1_HI, TILE + C_X1_LO, xb); put8(TILE + C_X1_HI + 8, TILE + C_X1_LO + 8, xb + 8);
        }
        // ---- G2: GRU pre-activations of both tiles (A = [x1 | h], K = 32)
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        if(warp == 0 && tc::elect_one()){
            tc::tc_fence_after();
#pragma unroll
            for(int e = 0; e < 2; e++){
                issue_gemm(e * TILE + C_D2, e * TILE + C_X1_HI, e * TILE + C_X1_LO, 2, TcImage::B2_HI, TcImage::B2_LO, 0, 64, IDESC64, 0);
                issue_gemm(e * TILE + C_D2, e * TILE + C_H_HI, e * TILE + C_H_LO, 2, TcImage::B2_HI, TcImage::B2_LO, 2, 64, IDESC64, 1);
            }
            tc::mma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, phase); phase ^= 1;
        tc::tc_fence_after();
        // ---- gate epilogue + dense 2 per environment (packed over two hidden units), new hidden state back to TMEM
        F2 act[4];
        auto gru = [&](auto E, int& gs){
            constexpr int e = decltype(E)::value;
            const uint32_t tb = lane_addr + e * TILE;
            const float* bias = sm_b + TcImage::BIAS2;
            const float* w2 = sm_b + TcImage::W2T;
            const float4 b2 = *reinterpret_cast<const float4*>(w2 + 64);
            F2 a01 = mk(b2.x, b2.y), a23 = mk(b2.z, b2.w);
            const int new_step = gs + 1;
            const bool wrap = !no_auto_reset && new_step >= a.seq_len;   // gru/operations_generic.h:400-410: the output is kept, the stored state resets
            const F2 one2 = bc(1.0f), minus2 = bc(-2.0f);
#pragma unroll
            for(int c = 0; c < 2; c++){
                float r[8], nx[8], nh[8], z[8], hh[8], hl[8], hn[8];
                tc::tmem_ld8(tb + C_D2 + 8 * c, r);
                tc::tmem_ld8(tb + C_D2 + 32 + 8 * c, nx);
                tc::tmem_ld8(tb + C_D2 + 48 + 8 * c, nh);
                tc::tmem_ld8(tb + C_D2 + 16 + 8 * c, z);
                tc::tmem_ld8(tb + C_H_HI + 8 * c, hh);
                tc::tmem_ld8(tb + C_H_LO + 8 * c, hl);
                tc::tmem_ld_wait();
#pragma unroll
                for(int j = 0; j < 8; j += 2){
                    const int u = 8 * c + j;
                    const F2 b_r = *reinterpret_cast<const F2*>(bias + u), b_z = *reinterpret_cast<const F2*>(bias + 16 + u);
                    const F2 b_x = *reinterpret_cast<const F2*>(bias + 32 + u), b_h = *reinterpret_cast<const F2*>(bias + 48 + u);
                    const F2 tr = add(mk(r[j], r[j + 1]), b_r);
                    const F2 dr = add(one2, mk(ex2_approx(tr.x), ex2_approx(tr.y)));
                    const F2 sg = mk(rcp_approx(dr.x), rcp_approx(dr.y));
                    const F2 tn = fma(add(mk(nh[j], nh[j + 1]), b_h), sg, add(mk(nx[j], nx[j + 1]), b_x));
                    const F2 dn = add(mk(ex2_approx(tn.x), ex2_approx(tn.y)), one2);
                    const F2 nn = fma(minus2, mk(rcp_approx(dn.x), rcp_approx(dn.y)), one2);
                    const F2 tz = add(mk(z[j], z[j + 1]), b_z);
                    const F2 dz = add(one2, mk(ex2_approx(tz.x), ex2_approx(tz.y)));
                    const F2 zz = mk(rcp_approx(dz.x), rcp_approx(dz.y));
                    const F2 h2 = add(mk(hh[j], hh[j + 1]), mk(hl[j], hl[j + 1]));
                    const F2 o = fma(zz, sub(h2, nn), nn);   // (1 - z) n + z h
                    hn[j] = o.x; hn[j + 1] = o.y;
                    const float4 wa = *reinterpret_cast<const float4*>(w2 + 4 * u), wb = *reinterpret_cast<const float4*>(w2 + 4 * u + 4);
                    a01 = fma(mk(wa.x, wa.y), bc(o.x), a01); a23 = fma(mk(wa.z, wa.w), bc(o.x), a23);
                    a01 = fma(mk(wb.x, wb.y), bc(o.y), a01); a23 = fma(mk(wb.z, wb.w), bc(o.y), a23);
                }
                if(wrap) put8(e * TILE + C_H_HI + 8 * c, e * TILE + C_H_LO + 8 * c, sm_b + TcImage::H0 + 8 * c);
                else put8(e * TILE + C_H_HI + 8 * c, e * TILE + C_H_LO + 8 * c, hn);
            }
            gs = wrap ? 0 : new_step;
            if(a.out_actions && active_of(E)) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env_of(E)) * 4) = make_float4(a01.x, a01.y, a23.x, a23.y);
            if constexpr(e == 0){ act[0].x = a01.x; act[1].x = a01.y; act[2].x = a23.x; act[3].x = a23.y; }
            else{ act[0].y = a01.x; act[1].y = a01.y; act[2].y = a23.x; act[3].y = a23.y; }
        };
        gru(E0{}, gs0); gru(E1{}, gs1);

        // ---- reward terms of the state BEFORE the step (squared/operations_generic.h:13-46; zero-weight terms are skipped, the weights are launch constants)
        const float4 la0 = sm_last[0], la1 = sm_last[BLOCK];
        F2 dpos[3], dvel[3];
        {
            const float4 u0 = sm_lang[0], u1 = sm_lang[BLOCK], v0 = sm_lang[3 * BLOCK], v1 = sm_lang[4 * BLOCK];   // position[3] velocity[3] ... of A (k = 0, 1) and B
            dpos[0] = mk(lang0 ? u0.x : 0.0f, lang1 ? v0.x : 0.0f); dpos[1] = mk(lang0 ? u0.y : 0.0f, lang1 ? v0.y : 0.0f); dpos[2] = mk(lang0 ? u0.z : 0.0f, lang1 ? v0.z : 0.0f);
            dvel[0] = mk(lang0 ? u0.w : 0.0f, lang1 ? v0.w : 0.0f); dvel[1] = mk(lang0 ? u1.x : 0.0f, lang1 ? v1.x : 0.0f); dvel[2] = mk(lang0 ? u1.y : 0.0f, lang1 ? v1.y : 0.0f);
        }
        auto sqrt2 = [](F2 v){ return mk(sqrt_approx(v.x), sqrt_approx(v.y)); };
        auto norm3 = [&](F2 u, F2 v, F2 w){ return sqrt2(fma(w, w, fma(v, v, mul(u, u)))); };
        F2 weighted = bc(0.0f), t_action = bc(0.0f), t_daction = bc(0.0f);
        if(const float w = R0[P_RW_POSITION]; w != 0.0f){
            F2 c = norm3(sub(x[X_POS], dpos[0]), sub(x[X_POS + 1], dpos[1]), sub(x[X_POS + 2], dpos[2]));
            const float clip = R0[P_RW_POSITION_CLIP];
            if(clip > 0.0f) c = mk(fminf(c.x, clip), fminf(c.y, clip));
            weighted = fma(bc(w), c, weighted);
        }
        if(const float w = R0[P_RW_ORIENTATION]; w != 0.0f){
            const F2 ac = acos_1m_x2(abs2(x[X_ORI + 3]));
            weighted = fma(bc(w), add(ac, ac), weighted);
        }
        if(const float w = R0[P_RW_LINVEL]; w != 0.0f) weighted = fma(bc(w), norm3(sub(x[X_VEL], dvel[0]), sub(x[X_VEL + 1], dvel[1]), sub(x[X_VEL + 2], dvel[2])), weighted);
        if(const float w = R0[P_RW_ANGVEL]; w != 0.0f) weighted = fma(bc(w), norm3(x[X_OMEGA], x[X_OMEGA + 1], x[X_OMEGA + 2]), weighted);
        if(const float w = R0[P_RW_ACTION]; w != 0.0f){
            const F2 hover = bc(R0[P_HOVER]), half = bc(0.5f);
            F2 acc = bc(0.0f);
#pragma unroll
            for(int i = 0; i < 4; i++){ const F2 dd = sub(fma(act[i], half, half), hover); acc = fma(dd, dd, acc); }
            const F2 c = sqrt2(acc);
            t_action = mul(bc(w), mul(c, c));
        }
        if(const float w = R0[P_RW_DACTION]; w != 0.0f){
            const F2 l[4] = {mk(la0.x, la1.x), mk(la0.y, la1.y), mk(la0.z, la1.z), mk(la0.w, la1.w)};
            F2 acc = bc(0.0f);
#pragma unroll
            for(int i = 0; i < 4; i++){ const F2 dd = sub(act[i], l[i]); acc = fma(dd, dd, acc); }
            t_daction = mul(bc(w), sqrt2(acc));
        }

        // ---- step: action scaling + RK4 + post integration (operations_generic.h:94-130, integrators.h:18-50, 70_post_integration.h)
        F2 xn[X_DIM];
        {
            F2 sp[4];
            const float4 hr4 = blk.q(X2_TA + 2), so4 = blk.q(X2_SP_OFFSET);   // {ta2, half_range}, {sp_offset, pad}
            const F2 hr = hi2(hr4), so = lo2(so4);
#pragma unroll
            for(int i = 0; i < 4; i++){
                const F2 c = mk(fminf(fmaxf(act[i].x, -1.0f), 1.0f), fminf(fmaxf(act[i].y, -1.0f), 1.0f));
                sp[i] = fma(c, hr, so);
            }
            const float4 t0 = blk.q(X2_DT), t1 = blk.q(X2_DT3);
            const F2 c1 = lo2(t0), c2 = hi2(t0), c3 = lo2(t1), c6 = hi2(t1);
            F2 k[X_DIM], tmp[X_DIM];
            dynamics_x2(blk, x, sp, k);
#pragma unroll
            for(int i = 0; i < X_DIM; i++){ xn[i] = fma(c6, k[i], x[i]); tmp[i] = fma(c2, k[i], x[i]); }
            dynamics_x2(blk, tmp, sp, k);
#pragma unroll
            for(int i = 0; i < X_DIM; i++){ xn[i] = fma(c3, k[i], xn[i]); tmp[i] = fma(c2, k[i], x[i]); }
            dynamics_x2(blk, tmp, sp, k);
#pragma unroll
            for(int i = 0; i < X_DIM; i++){ xn[i] = fma(c3, k[i], xn[i]); tmp[i] = fma(c1, k[i], x[i]); }
            dynamics_x2(blk, tmp, sp, k);
#pragma unroll
            for(int i = 0; i < X_DIM; i++) xn[i] = fma(c6, k[i], xn[i]);
        }
        {
            const F2 q0 = xn[X_ORI], q1 = xn[X_ORI + 1], q2 = xn[X_ORI + 2], q3 = xn[X_ORI + 3];
            const F2 nrm = fma(q3, q3, fma(q2, q2, fma(q1, q1, mul(q0, q0))));
            const F2 inv = mk(rsqrt_approx(nrm.x), rsqrt_approx(nrm.y));
#pragma unroll
            for(int i = 0; i < 4; i++) xn[X_ORI + i] = mul(xn[X_ORI + i], inv);
            // clamp of position / velocities to +-1e5 (70_post_integration.h:29-36): one range test per environment, the clamps only when it fails
            auto guard = [&](auto E){
                constexpr int e = decltype(E)::value;
                const float m = max3(max3(fabsf(lane<e>(xn[0])), fabsf(lane<e>(xn[1])), fabsf(lane<e>(xn[2]))),
                                     max3(fabsf(lane<e>(xn[7])), fabsf(lane<e>(xn[8])), fabsf(lane<e>(xn[9]))),
                                     max3(fabsf(lane<e>(xn[10])), fabsf(lane<e>(xn[11])), fabsf(lane<e>(xn[12]))));
                return !(m <= 100000.0f);   // also true for NaN
            };
            if(guard(E0{}) || guard(E1{})){
#pragma unroll
                for(int i = 0; i < 3; i++){
                    F2& p_ = xn[X_POS + i]; F2& v_ = xn[X_VEL + i]; F2& w_ = xn[X_OMEGA + i];
                    p_ = mk(clamp_t<true>(p_.x, -100000.0f, 100000.0f), clamp_t<true>(p_.y, -100000.0f, 100000.0f));
                    v_ = mk(clamp_t<true>(v_.x, -100000.0f, 100000.0f), clamp_t<true>(v_.y, -100000.0f, 100000.0f));
                    w_ = mk(clamp_t<true>(w_.x, -100000.0f, 100000.0f), clamp_t<true>(w_.y, -100000.0f, 100000.0f));
                }
            }
            const F2 amin = hi2(blk.q(X2_JD + 2)), amax = hi2(blk.q(X2_JID + 2));
#pragma unroll
            for(int i = 0; i < 4; i++){
                F2& r_ = xn[X_RPM + i];
                r_ = mk(fminf(fmaxf(r_.x, amin.x), amax.x), fminf(fmaxf(r_.y, amin.y), amax.y));
            }
        }
        // ---- terminated (operations_generic.h:142-166) of the new state
        bool term0 = false, term1 = false;
        if(R0[P_TERM_ENABLED] != 0.0f){
            const float tv = R0[P_TERM_LINVEL], tw = R0[P_TERM_ANGVEL];
            const F2 tp = blk.p(X2_TERM_POS);
            auto test = [&](auto E, float tpe){
                constexpr int e = decltype(E)::value;
                return max3(fabsf(lane<e>(xn[0])), fabsf(lane<e>(xn[1])), fabsf(lane<e>(xn[2]))) > tpe ||
                       max3(fabsf(lane<e>(xn[7])), fabsf(lane<e>(xn[8])), fabsf(lane<e>(xn[9]))) > tv ||
                       max3(fabsf(lane<e>(xn[10])), fabsf(lane<e>(xn[11])), fabsf(lane<e>(xn[12]))) > tw;
            };
            term0 = test(E0{}, tp.x); term1 = test(E1{}, tp.y);
        }
        // ---- remaining reward terms (need the new velocities), total, flags
        {
            const F2 dt = blk.p(X2_DT);
            if(const float w = R0[P_RW_LINACC]; w != 0.0f){
                const F2 c = norm3(sub(xn[X_VEL], x[X_VEL]), sub(xn[X_VEL + 1], x[X_VEL + 1]), sub(xn[X_VEL + 2], x[X_VEL + 2]));
                weighted = fma(bc(w), mk(c.x / dt.x, c.y / dt.y), weighted);
            }
            if(const float w = R0[P_RW_ANGACC]; w != 0.0f){
                const F2 c = norm3(sub(xn[X_OMEGA], x[X_OMEGA]), sub(xn[X_OMEGA + 1], x[X_OMEGA + 1]), sub(xn[X_OMEGA + 2], x[X_OMEGA + 2]));
                weighted = fma(bc(w), mk(c.x / dt.x, c.y / dt.y), weighted);
            }
            weighted = add(add(weighted, t_action), t_daction);
            const F2 free_ = fnma(bc(R0[P_RW_SCALE]), weighted, bc(R0[P_RW_CONSTANT]));   // -scale * weighted + constant
            const bool nonneg = R0[P_RW_NONNEG] != 0.0f;
            const float pen = R0[P_RW_TERM_PENALTY];
            const float r0 = term0 ? pen : ((free_.x > 0.0f || !nonneg) ? free_.x : 0.0f);
            const float r1 = term1 ? pen : ((free_.y > 0.0f || !nonneg) ? free_.y : 0.0f);
            if(a.out_rewards){ if(act0) a.out_rewards[(size_t)t * n + env0] = r0; if(act1) a.out_rewards[(size_t)t * n + env1] = r1; }
            if(a.out_term){ if(act0) a.out_term[(size_t)t * n + env0] = term0 ? 1 : 0; if(act1) a.out_term[(size_t)t * n + env1] = term1 ? 1 : 0; }
            if(!done0){ ret.x += r0; eplen0 += 1; done0 = term0; }
            if(!done1){ ret.y += r1; eplen1 += 1; done1 = term1; }
        }
        // ---- commit the new state: last action, action history (H = 1), Langevin target (70_post_integration.h:40-48, :112-125, :127-170)
#pragma unroll
        for(int i = 0; i < X_DIM; i++) x[i] = xn[i];
#pragma unroll
        for(int i = 0; i < 4; i++) hist[i] = act[i];
        sm_last[0] = make_float4(act[0].x, act[1].x, act[2].x, act[3].x);
        sm_last[BLOCK] = make_float4(act[0].y, act[1].y, act[2].y, act[3].y);
        auto langevin = [&](auto E, uint64_t& rng){
            constexpr int e = decltype(E)::value;
            const float gamma = R0[P_LANGEVIN_GAMMA], omega = R0[P_LANGEVIN_OMEGA], sigma = R0[P_LANGEVIN_SIGMA], alpha = R0[P_LANGEVIN_ALPHA];
            const float dt = sm_row[2 * X2_DT + e], sqrt_dt = sm_row[2 * X2_SQRT_DT + e];
            const float4 v0 = sm_lang[(e * 3) * BLOCK], v1 = sm_lang[(e * 3 + 1) * BLOCK], v2 = sm_lang[(e * 3 + 2) * BLOCK];
            float L[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
            for(int dim = 0; dim < 3; dim++){
                const float x_prev = L[6 + dim], v_prev = L[9 + dim];
                const float dW = sqrt_dt * rng_normal_t<false, true>(rng, 0.0f, 1.0f);
                const float v_next = v_prev + (-gamma * v_prev - omega * omega * x_prev) * dt + sigma * dW;
                const float x_next = x_prev + v_next * dt;
                L[6 + dim] = x_next; L[9 + dim] = v_next;
                const float v_smooth = alpha * v_next + (1.0f - alpha) * L[3 + dim];
                L[dim] = L[dim] + v_smooth * dt;
                L[3 + dim] = v_smooth;
            }
            sm_lang[(e * 3) * BLOCK] = make_float4(L[0], L[1], L[2], L[3]);
            sm_lang[(e * 3 + 1) * BLOCK] = make_float4(L[4], L[5], L[6], L[7]);
            sm_lang[(e * 3 + 2) * BLOCK] = make_float4(L[8], L[9], L[10], L[11]);
        };
        if(lang0) langevin(E0{}, rng0);
        if(lang1) langevin(E1{}, rng1);
    }
    // ---- item end: state, hidden state, streams and accumulators back to HBM
    tc::tmem_st_wait();
    const bool last_chunk = chunk == n_chunks - 1;
    if(last_chunk && a.out_states && (a.T % a.state_stride) == 0){
        float* base = a.out_states + (size_t)(a.T / a.state_stride) * n * Spec::STATE_DIM;
        if(act0) write_row(E0{}, base + env0 * Spec::STATE_DIM);
        if(act1) write_row(E1{}, base + env1 * Spec::STATE_DIM);
    }
    auto store = [&](auto E, uint64_t rng, int gs, int eplen, bool done){
        constexpr int e = decltype(E)::value;
        float hh[HD], hl[HD];   // tcgen05.ld is warp-collective: every lane executes it, only active environments store
        tc::tmem_ld16(lane_addr + e * TILE + C_H_HI, hh);
        tc::tmem_ld16(lane_addr + e * TILE + C_H_LO, hl);
        tc::tmem_ld_wait();
        if(!active_of(E)) return;
        const size_t env = env_of(E);
        float* s = a.state + env;
#pragma unroll
        for(int i = 0; i < 13; i++) s[(size_t)i * n] = lane<e>(x[i]);
#pragma unroll
        for(int i = 0; i < 4; i++){ s[(size_t)(S_RPM + i) * n] = lane<e>(x[X_RPM + i]); s[(size_t)(S_HIST + i) * n] = lane<e>(hist[i]); }
        const float4 la = sm_last[e * BLOCK];
        s[(size_t)(S_LAST_ACTION + 0) * n] = la.x; s[(size_t)(S_LAST_ACTION + 1) * n] = la.y; s[(size_t)(S_LAST_ACTION + 2) * n] = la.z; s[(size_t)(S_LAST_ACTION + 3) * n] = la.w;
#pragma unroll
        for(int i = 0; i < 3; i++) s[(size_t)(S_ANGVEL_HIST + i) * n] = lane<e>(x[X_OMEGA + i]);   // history length 0: copy of omega (70_post_integration.h:63-67)
        s[(size_t)S_CURRENT_STEP * n] = 0.0f;                                                       // (0 + 1) % 1
#pragma unroll
        for(int k = 0; k < 3; k++){
            const float4 v = sm_lang[(e * 3 + k) * BLOCK];
            const int r = s_langevin(1) + 4 * k;
            s[(size_t)r * n] = v.x; s[(size_t)(r + 1) * n] = v.y; s[(size_t)(r + 2) * n] = v.z; s[(size_t)(r + 3) * n] = v.w;
        }
        a.rng[env] = rng;
#pragma unroll
        for(int j = 0; j < HD; j++) a.hidden[(size_t)j * n + env] = hh[j] + hl[j];
        a.gru_step[env] = gs;
        if(last_chunk){
            if(a.out_returns) a.out_returns[env] = lane<e>(ret);
            if(a.out_eplen) a.out_eplen[env] = eplen;
            if(a.out_done) a.out_done[env] = done ? 1 : 0;
        }
        else{ a.acc_ret[env] = lane<e>(ret); a.acc_len[env] = (eplen << 1) | (done ? 1 : 0); }
    };
    store(E0{}, rng0, gs0, eplen0, done0); store(E1{}, rng1, gs1, eplen1, done1);
    if(!last_chunk){   // publish: every thread's stores, then the pair's progress counter
        __threadfence();
        __syncthreads();
        if(tid == 0) atomicExch(a.sched + 1 + pair, chunk + 1);
    }
    }   // work loop (the staged rows, parked vectors and TMEM lanes are private to their thread: no barrier needed before the next item restages them)
    tc::tc_fence_before();
    __syncthreads();
    if(warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

}  // namespace b200l2f
