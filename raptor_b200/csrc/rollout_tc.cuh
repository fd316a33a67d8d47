// raptor_b200/csrc/rollout_tc.cuh -- fused persistent rollout with the actor GEMMs on the 5th-generation tensor cores.
//
// One CTA = 128 threads = 128 environments = the M dimension of every MMA.  Per control step the CTA runs three dependent GEMMs
//   G1  [128 x 24] x [24 x 16]   obs(22) | 1 | 0          -> dense-1 pre-activation (bias folded in as a K column)
//   G2  [128 x 40] x [40 x 64]   x1(16) | h(16) | 1 1 0.. -> r,z pre-activations (32), n_x (16), n_h (16)   (block-structured B)
//   G3  [128 x 24] x [24 x 16]   h'(16) | 1 | 0           -> action (4 of 16 columns used)
// as tcgen05.mma.kind::tf32 with the 3xTF32 error-compensated split (A_hi B_hi + A_lo B_hi + A_hi B_lo; plain TF32 has a 10-bit
// mantissa and breaks the 1e-4 closed-loop parity bound).  Activations are produced on chip: every thread writes ITS row of the A
// operand (hi and lo planes) straight into the canonical K-major core-matrix layout in shared memory with conflict-free STS.128,
// one elected thread issues the MMAs, accumulators live in TMEM and each thread reads back its own row with tcgen05.ld (lane = env).
// The weight image (hi/lo planes of the three B operands, 26 KB) arrives once per CTA by a 1-D TMA bulk copy.
// Everything that is not a dense contraction (observe, gates, RK4, reward) stays on the fp32 CUDA cores exactly as in k_rollout_raptor,
// except that the per-rotor force/torque accumulation uses the per-environment rotor matrices A_F, A_T (thrust and torque are linear in
// the four rotor thrusts), which shrinks the staged parameter block from 86 to 67 floats and a dynamics evaluation by ~50 instructions.
#pragma once
#include "kernels.cuh"
#include "tc.cuh"

namespace b200l2f {

// ---- B operand image (host-built): for each GEMM two planes (hi, lo), each [K/4][N][4] floats ------------------------------
struct TcImage {
    static constexpr int K1 = 24, N1 = 16, K2 = 40, N2 = 64, K3 = 24, N3 = 16;
    static constexpr int B1_HI = 0, B1_LO = B1_HI + K1 * N1;
    static constexpr int B2_HI = B1_LO + K1 * N1, B2_LO = B2_HI + K2 * N2;
    static constexpr int B3_HI = B2_LO + K2 * N2, B3_LO = B3_HI + K3 * N3;
    static constexpr int H0 = B3_LO + K3 * N3;   // initial hidden state (16 floats) for the auto-reset
    static constexpr int W2T = H0 + 16;          // fp32 dense-2 weights, k-major [16][4], then b2[4]
    static constexpr int W1T = W2T + 68;         // fp32 dense-1 weights, k-major [22][16], then b1[16]
    static constexpr int BIAS2 = W1T + 22 * 16 + 16;  // fp32 GRU biases for the epilogue of the TMEM-A variant: (b_ih + b_hh)[0:32] | b_in[16] | b_hn[16]
    static constexpr int SIZE = BIAS2 + 64;      // floats
    static constexpr int BYTES = SIZE * 4;
    static_assert(BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
};
inline void tc_split_host(float x, float& hi, float& lo){
    uint32_t u; std::memcpy(&u, &x, 4); u &= 0xFFFFE000u; std::memcpy(&hi, &u, 4); lo = x - hi;
}
// blob (include/b200_l2f.h RAPTOR_GRU order) -> image.  scaled_gates (TMEM-A kernel): the GRU rows and biases carry the exponent scale of
// their activation -- r, z rows times -log2(e), n rows times 2 log2(e) -- so the epilogue feeds the accumulators straight into ex2.
inline void build_tc_image_host(float* img, const float* blob, bool scaled_gates = false){
    const float s_rz = scaled_gates ? -1.4426950408889634f : 1.0f, s_n = scaled_gates ? 2.8853900817779268f : 1.0f;
    const float s_x = scaled_gates ? 0.5f : 1.0f;   // TMEM-A kernel: dense 1's ReLU arrives doubled (x1 + |x1| on the packed pipe); the exact factor 0.5 sits in the W_ih columns
    constexpr int IN = 22, HD = 16, OUT = 4;
    const float* W1 = blob; const float* b1 = W1 + HD * IN;
    const float* Wih = b1 + HD; const float* bih = Wih + 3 * HD * HD;
    const float* Whh = bih + 3 * HD; const float* bhh = Whh + 3 * HD * HD;
    const float* h0 = bhh + 3 * HD; const float* W2 = h0 + HD; const float* b2 = W2 + OUT * HD;
    for(int i = 0; i < TcImage::SIZE; i++) img[i] = 0.0f;
    auto put = [&](int hi_base, int lo_base, int N, int n, int k, float v){
        float hi, lo; tc_split_host(v, hi, lo);
        const int idx = (k / 4) * N * 4 + n * 4 + (k % 4);
        img[hi_base + idx] = hi; img[lo_base + idx] = lo;
    };
    for(int n = 0; n < HD; n++){
        for(int k = 0; k < IN; k++) put(TcImage::B1_HI, TcImage::B1_LO, 16, n, k, W1[n * IN + k]);
        put(TcImage::B1_HI, TcImage::B1_LO, 16, n, 22, b1[n]);
    }
    // G2 columns: 0..15 x1, 16..31 h, 32 -> b_hh, 33 -> b_ih (A holds 1.0 in both)
    for(int j = 0; j < 2 * HD; j++){            // r, z rows
        for(int k = 0; k < HD; k++){ put(TcImage::B2_HI, TcImage::B2_LO, 64, j, k, (s_rz * Wih[j * HD + k]) * s_x); put(TcImage::B2_HI, TcImage::B2_LO, 64, j, 16 + k, s_rz * Whh[j * HD + k]); }
        put(TcImage::B2_HI, TcImage::B2_LO, 64, j, 32, s_rz * bhh[j]); put(TcImage::B2_HI, TcImage::B2_LO, 64, j, 33, s_rz * bih[j]);
    }
    for(int j = 0; j < HD; j++){                // n_x rows 32..47, n_h rows 48..63
        for(int k = 0; k < HD; k++){ put(TcImage::B2_HI, TcImage::B2_LO, 64, 32 + j, k, (s_n * Wih[(2 * HD + j) * HD + k]) * s_x); put(TcImage::B2_HI, TcImage::B2_LO, 64, 48 + j, 16 + k, s_n * Whh[(2 * HD + j) * HD + k]); }
        put(TcImage::B2_HI, TcImage::B2_LO, 64, 32 + j, 33, s_n * bih[2 * HD + j]);
        put(TcImage::B2_HI, TcImage::B2_LO, 64, 48 + j, 32, s_n * bhh[2 * HD + j]);
    }
    for(int n = 0; n < OUT; n++){
        for(int k = 0; k < HD; k++) put(TcImage::B3_HI, TcImage::B3_LO, 16, n, k, W2[n * HD + k]);
        put(TcImage::B3_HI, TcImage::B3_LO, 16, n, 16, b2[n]);
    }
    for(int j = 0; j < HD; j++) img[TcImage::H0 + j] = h0[j];
    for(int k = 0; k < HD; k++) for(int n = 0; n < OUT; n++) img[TcImage::W2T + 4 * k + n] = W2[n * HD + k];
    for(int n = 0; n < OUT; n++) img[TcImage::W2T + 64 + n] = b2[n];
    for(int k = 0; k < IN; k++) for(int n = 0; n < HD; n++) img[TcImage::W1T + k * HD + n] = W1[n * IN + k];
    for(int n = 0; n < HD; n++) img[TcImage::W1T + IN * HD + n] = b1[n];
    for(int j = 0; j < 2 * HD; j++) img[TcImage::BIAS2 + j] = s_rz * (bih[j] + bhh[j]);
    for(int j = 0; j < HD; j++){ img[TcImage::BIAS2 + 32 + j] = s_n * bih[2 * HD + j]; img[TcImage::BIAS2 + 48 + j] = s_n * bhh[2 * HD + j]; }
}

#ifndef B200L2F_G1_CUDA
#define B200L2F_G1_CUDA 0          // dense 1 of k_rollout_raptor_ts on the packed CUDA-core pipe instead of a tensor-core round trip: experiment, see start_g1
#endif
#ifndef B200L2F_GATES_PACKED
#define B200L2F_GATES_PACKED 1     // GRU gate epilogue of k_rollout_raptor_ts on the packed fp32 pipe (two hidden units per FADD2 / FFMA2)
#endif
#ifndef B200L2F_TMEM_PREFETCH
#define B200L2F_TMEM_PREFETCH 0    // issue the second batch of accumulator loads (z, h) before the MUFU work on the first (r, n); measured -0.4 %
#endif
#ifndef B200L2F_HOIST_LANGEVIN
#define B200L2F_HOIST_LANGEVIN 0   // draw the Langevin target's normals in the shadow of the first MMA round trip (noise-free kernels only); measured -0.8 %
#endif
#ifndef B200L2F_SPLIT_PACKED
#define B200L2F_SPLIT_PACKED 0     // 3xTF32 operand split with the low parts on the packed pipe: lo = fma2(hi, -1, x); measured -1.3 % (register moves, 8 B of spills)
#endif
#ifndef B200L2F_SPLIT_PAIRS
#define B200L2F_SPLIT_PAIRS 1      // 3xTF32 low parts of x1 / h' (values that already sit in register pairs) as one FADD2 per two values
#endif
#ifndef B200L2F_DENSE2_PACKED
#define B200L2F_DENSE2_PACKED 1    // dense 2 (16 -> 4) as FFMA2 on (act0, act1) / (act2, act3); measured +1.2 %
#endif
#ifndef B200L2F_H_IN_SMEM
#define B200L2F_H_IN_SMEM 0        // the epilogue's copy of the GRU hidden state in shared memory (8 KB per CTA) instead of re-reading it from TMEM; measured -0.7 %
#endif
#ifndef B200L2F_PIPELINE_G1
#define B200L2F_PIPELINE_G1 0      // k_rollout_raptor_ts: dense 1 of step t + 1 issued before the reward / termination / Langevin work of step t; measured -4 % (profiles/r02_exp4_*)
#endif
#ifndef B200L2F_LANGEVIN_BRANCH_FREE
#define B200L2F_LANGEVIN_BRANCH_FREE 1   // default-math kernels: Langevin target update committed by selects instead of a branch (lets the scheduler interleave the RNG chain)
#endif
#ifndef B200L2F_PACKED_DYNAMICS
#define B200L2F_PACKED_DYNAMICS 1  // axial vehicles, default math: the dynamics evaluation itself on natural pairs (dynamics_axial_packed): 55 instead of ~110 instructions
#endif
#ifndef B200L2F_PACKED_FP32
#define B200L2F_PACKED_FP32 1      // packed fp32 (FFMA2 / FADD2 / FMUL2, sm_100) in the default-math kernels; 0 = scalar twins (tuning / bisecting)
#endif

// ---- compiled per-environment dynamics block (68 floats), staged thread-major as sm_dyn[tid * C_DIM + i] ------------------------------
// Every read is one LDS.128: the row stride is 68 words (= 4 mod 32), so the eight threads of a quarter-warp hit 8 x 4 distinct banks
// (conflict-free), and a dynamics evaluation issues ~10 loads instead of 37 scalar ones.  Groups are laid out as the float4s the axial
// vehicle's evaluation consumes; the general vehicle's extra entries (A_F, off-diagonal inertia) follow.
enum DynC : int { C_COEF = 0,        // [3][4] thrust-curve coefficients, k-major: coefficient k of rotor r at [4 k + r] (two rotors per packed operand)
                  C_AT01 = 12,       // [4][2] torque per unit rotor thrust, rows 0 and 1 of rotor r at [2 r], [2 r + 1] (one packed operand per rotor)
                  C_AT2 = 20,        // [4]    row 2
                  C_TAU_M = 24,      // [4] (1/tau_rise + 1/tau_fall) / 2
                  C_TAU_H = 28,      // [4] (1/tau_rise - 1/tau_fall) / 2: (setpoint - rpm) / tau(sign) = m d + h |d|, d = setpoint - rpm
                  C_G = 32,          // gravity[3] | action min
                  C_JD = 36,         // diag(J)[3] | action max
                  C_JID = 40,        // diag(J^-1)[3] | termination position threshold
                  C_DT = 44,         // dt | dt / 2 | dt / 3 | dt / 6   (the RK4 weights of integrators.h:18-50, divided once per staging)
                  C_SQRT_DT = 48,    // sqrt(dt) | 0 | 0 | 0            (Langevin target, 70_post_integration.h:135)
                  C_AF = 52,         // [3][4] force per unit rotor thrust (general vehicle only)
                  C_JOFF = 64, C_JIOFF = 70,   // off-diagonal entries of J, J^-1 in row-major order 01 02 10 12 20 21 (general vehicle only)
                  C_DIM_AXIAL = 52,  // the axial vehicle's evaluation reads nothing beyond this: a compact block for kernels that need the shared memory
                  C_DIM = 76 };
static_assert(C_DIM_AXIAL % 4 == 0 && (C_DIM_AXIAL / 4) % 2 == 1, "row stride = odd number of 16-byte units: a quarter-warp's LDS.128 hits 8 x 4 distinct banks");
static_assert(C_DIM % 4 == 0 && (C_DIM / 4) % 2 == 1, "row stride must keep LDS.128 aligned and conflict-free");
__host__ __device__ constexpr int dyn_j_slot(int base_diag, int base_off, int i, int j){   // where entry (i, j) of J / J^-1 lives in the block
    return i == j ? base_diag + i : base_off + 2 * i + (j > i ? j - 1 : j);
}
// NC: the parameter columns are read-only for the whole launch (ld.global.nc); false when the kernel itself rewrites them (collect's resets)
// FOLLOW: every entry outside the domain-randomised set equals row0 (collect, see k_collect_ts)
template <bool UNIFORM, bool NC = true, bool FOLLOW = false>
struct ParamsCompiledT {
    const float* sm; const float* base; size_t stride;                // sm: this thread's staged block; base/stride: full parameter column in HBM
    const float* row0;                                                // UNIFORM: environment 0's row in the launch's constant bank
    __device__ __forceinline__ float4 c4(int i) const { return *reinterpret_cast<const float4*>(sm + i); }      // i % 4 == 0
    __device__ __forceinline__ float c(int i) const {                 // compile-time i: one LDS.128 + a register pick (loads of a group are merged)
        const float4 v = c4(i & ~3);
        return (i & 3) == 0 ? v.x : (i & 3) == 1 ? v.y : (i & 3) == 2 ? v.z : v.w;
    }
    __device__ __forceinline__ float operator[](int i) const {       // everything outside the dynamics block
        if(i == P_TERM_POS) return c(C_JID + 3);                      // per environment even under DR (10_sample_initial_parameters.h:154)
        if(i == P_ACT_MIN) return c(C_G + 3);
        if(i == P_ACT_MAX) return c(C_JD + 3);
        if(UNIFORM && mdp_uniform_index(i)) return row0[i];
        if(FOLLOW && !dr_overlay_index(i)) return row0[i];
        return NC ? __ldg(base + (size_t)i * stride) : base[(size_t)i * stride];
    }
};
template <int STRIDE = C_DIM>
__device__ __forceinline__ float* dyn_block_of_thread(float* sm_dyn){ return sm_dyn + threadIdx.x * STRIDE; }
using ParamsCompiled = ParamsCompiledT<false>;
// compile this thread's dynamics block from any parameter accessor P(i) (HBM column, register overlay, ...)
// FAST: MUFU reciprocals / square root (the in-kernel resets of the default-math collection kernels; ~2 ulp on the rotor time constants and RK4 weights)
template <bool FULL = true, bool FAST = false, class F>
__device__ __forceinline__ void compile_dynamics_block(float* __restrict__ sm, F&& P){   // sm = dyn_block_of_thread(sm_dyn); FULL = false: axial entries only
#pragma unroll
    for(int r = 0; r < 4; r++){
#pragma unroll
        for(int k = 0; k < 3; k++) sm[C_COEF + 4 * k + r] = P(P_THRUST_COEF + 3 * r + k);
        const float dx = P(P_THRUST_DIR + 3 * r), dy = P(P_THRUST_DIR + 3 * r + 1), dz = P(P_THRUST_DIR + 3 * r + 2);
        const float px = P(P_ROTOR_POS + 3 * r), py = P(P_ROTOR_POS + 3 * r + 1), pz = P(P_ROTOR_POS + 3 * r + 2);
        const float kq = P(P_TORQUE_CONST + r);
        if constexpr(FULL){ sm[C_AF + 0 * 4 + r] = dx; sm[C_AF + 1 * 4 + r] = dy; sm[C_AF + 2 * 4 + r] = dz; }
        // torque of rotor r per unit thrust: torque_dir * k_q + r x dir   (60_dynamics.h:38-39)
        sm[C_AT01 + 2 * r + 0] = P(P_TORQUE_DIR + 3 * r + 0) * kq + (py * dz - pz * dy);
        sm[C_AT01 + 2 * r + 1] = P(P_TORQUE_DIR + 3 * r + 1) * kq + (pz * dx - px * dz);
        sm[C_AT2 + r]          = P(P_TORQUE_DIR + 3 * r + 2) * kq + (px * dy - py * dx);
        const float ir = FAST ? rcp_approx(P(P_TAU_RISE + r)) : 1.0f / P(P_TAU_RISE + r), ifl = FAST ? rcp_approx(P(P_TAU_FALL + r)) : 1.0f / P(P_TAU_FALL + r);
        sm[C_TAU_M + r] = 0.5f * (ir + ifl);
        sm[C_TAU_H + r] = 0.5f * (ir - ifl);
    }
#pragma unroll
    for(int i = 0; i < 3; i++) sm[C_G + i] = P(P_GRAVITY + i);
#pragma unroll
    for(int i = 0; i < 3; i++){
#pragma unroll
        for(int j = 0; j < 3; j++){
            if(FULL || i == j){
                sm[dyn_j_slot(C_JD, C_JOFF, i, j)] = P(P_J + 3 * i + j);
                sm[dyn_j_slot(C_JID, C_JIOFF, i, j)] = P(P_JINV + 3 * i + j);
            }
        }
    }
    sm[C_G + 3] = P(P_ACT_MIN); sm[C_JD + 3] = P(P_ACT_MAX); sm[C_JID + 3] = P(P_TERM_POS);
    const float dt = P(P_DT);
    sm[C_DT] = dt; sm[C_DT + 1] = dt / 2.0f; sm[C_DT + 2] = FAST ? dt * 0.333333343267440796f : dt / 3.0f; sm[C_DT + 3] = FAST ? dt * 0.16666667163372040f : dt / 6.0f;
    sm[C_SQRT_DT] = sqrt_t<FAST>(dt); sm[C_SQRT_DT + 1] = 0.0f; sm[C_SQRT_DT + 2] = 0.0f; sm[C_SQRT_DT + 3] = 0.0f;
}
template <bool UNIFORM, bool NC = true, bool FOLLOW = false, int STRIDE = C_DIM>
__device__ __forceinline__ ParamsCompiledT<UNIFORM, NC, FOLLOW> stage_dynamics_compiled(float* __restrict__ sm_dyn, const float* params, size_t n, size_t env, const float* row0){
    float* sm = dyn_block_of_thread<STRIDE>(sm_dyn);
    const float* g = params + env;
    compile_dynamics_block<STRIDE == C_DIM>(sm, [&](int i){ return NC ? __ldg(g + (size_t)i * n) : g[(size_t)i * n]; });
    ParamsCompiledT<UNIFORM, NC, FOLLOW> p; p.sm = sm; p.base = g; p.stride = n; p.row0 = row0;
    return p;
}
// ---- packed fp32 helpers (sm_100: FFMA2 / FADD2 / FMUL2 execute one operation on two fp32 lanes per instruction; operands may be a register pair,
// ---- a broadcast scalar register, a uniform register or an immediate, each with negate / absolute-value modifiers) -------------------------------
using F2 = float2;
namespace p2 {
__device__ __forceinline__ F2 mk(float a, float b){ return make_float2(a, b); }
__device__ __forceinline__ F2 bc(float s){ return make_float2(s, s); }                       // SASS: a .F32 (broadcast) operand, no move
__device__ __forceinline__ F2 neg(F2 a){ return make_float2(-a.x, -a.y); }                   // operand modifier
__device__ __forceinline__ F2 abs2(F2 a){ return make_float2(fabsf(a.x), fabsf(a.y)); }      // operand modifier
__device__ __forceinline__ F2 add(F2 a, F2 b){ return __fadd2_rn(a, b); }
__device__ __forceinline__ F2 sub(F2 a, F2 b){ return __fadd2_rn(a, neg(b)); }
__device__ __forceinline__ F2 mul(F2 a, F2 b){ return __fmul2_rn(a, b); }
__device__ __forceinline__ F2 fma(F2 a, F2 b, F2 c){ return __ffma2_rn(a, b, c); }
__device__ __forceinline__ F2 fnma(F2 a, F2 b, F2 c){ return __ffma2_rn(neg(a), b, c); }    // c - a b
template <int E> __device__ __forceinline__ float lane(F2 v){ return E == 0 ? v.x : v.y; }
__device__ __forceinline__ F2 lo2(float4 v){ return make_float2(v.x, v.y); }
__device__ __forceinline__ F2 hi2(float4 v){ return make_float2(v.z, v.w); }
using b200l2f::max3;
}  // namespace p2

// multirotor dynamics with the rotor matrices (same physics as dynamics() in env.cuh; thrust/torque summed as matrix-vector products).
// AXIAL: every rotor thrusts along body z and J, J^-1 are diagonal (all reference vehicles; k_param_features bit2 guards it): the zero
// products are dropped, the remaining operations keep the order of the general form, so both forms give the same bits.
template <bool AXIAL = false, class PC>
__device__ __forceinline__ void dynamics_compiled(const PC& p, const DynInvariants& d, const float* __restrict__ x, const float* __restrict__ setpoint, float* __restrict__ dx){
    float tm[4];
    {
        const float4 k0 = p.c4(C_COEF), k1 = p.c4(C_COEF + 4), k2 = p.c4(C_COEF + 8);
        const float c0[4] = {k0.x, k0.y, k0.z, k0.w}, c1[4] = {k1.x, k1.y, k1.z, k1.w}, c2[4] = {k2.x, k2.y, k2.z, k2.w};
#pragma unroll
        for(int r = 0; r < 4; r++){
            const float rpm = x[X_RPM + r];
            tm[r] = c0[r] + c1[r] * rpm + c2[r] * rpm * rpm;
        }
    }
    float thrust[3], torque[3];
#pragma unroll
    for(int i = 0; i < 3; i++){
        if constexpr(!AXIAL){ const float4 f = p.c4(C_AF + 4 * i); thrust[i] = f.x * tm[0] + f.y * tm[1] + f.z * tm[2] + f.w * tm[3]; }
    }
    {
        const float4 u = p.c4(C_AT01), v = p.c4(C_AT01 + 4), w = p.c4(C_AT2);
        torque[0] = u.x * tm[0] + u.z * tm[1] + v.x * tm[2] + v.z * tm[3];
        torque[1] = u.y * tm[0] + u.w * tm[1] + v.y * tm[2] + v.w * tm[3];
        torque[2] = w.x * tm[0] + w.y * tm[1] + w.z * tm[2] + w.w * tm[3];
    }
    if constexpr(AXIAL) thrust[2] = ((tm[0] + tm[1]) + tm[2]) + tm[3];
#pragma unroll
    for(int i = 0; i < 3; i++) dx[X_POS + i] = x[X_VEL + i];
    const float q0 = x[X_ORI], q1 = x[X_ORI + 1], q2 = x[X_ORI + 2], q3 = x[X_ORI + 3];
    const float w0 = x[X_OMEGA], w1 = x[X_OMEGA + 1], w2 = x[X_OMEGA + 2];
    dx[X_ORI + 0] = (-q1 * w0 - q2 * w1 - q3 * w2) * 0.5f;
    dx[X_ORI + 1] = ( q0 * w0 + q2 * w2 - q3 * w1) * 0.5f;
    dx[X_ORI + 2] = ( q0 * w1 + q3 * w0 - q1 * w2) * 0.5f;
    dx[X_ORI + 3] = ( q0 * w2 + q1 * w1 - q2 * w0) * 0.5f;
    {
        float o0, o1, o2;
        if constexpr(AXIAL){   // rotate (0, 0, T) by q
            const float T = thrust[2];
            const float v0 = (q2 * T) * 2.0f, v1 = (-(q1 * T)) * 2.0f;
            o0 = -(q3 * v1); o1 = q3 * v0; o2 = q1 * v1 - q2 * v0;
            o0 += v0 * q0; o1 += v1 * q0;
            o2 += T;
        }
        else{
            float v0 = (q2 * thrust[2] - q3 * thrust[1]) * 2.0f;
            float v1 = (q3 * thrust[0] - q1 * thrust[2]) * 2.0f;
            float v2 = (q1 * thrust[1] - q2 * thrust[0]) * 2.0f;
            o0 = q2 * v2 - q3 * v1; o1 = q3 * v0 - q1 * v2; o2 = q1 * v1 - q2 * v0;
            o0 += v0 * q0; o1 += v1 * q0; o2 += v2 * q0;
            o0 += thrust[0]; o1 += thrust[1]; o2 += thrust[2];
        }
        const float4 g = p.c4(C_G);
        dx[X_VEL + 0] = o0 * d.inv_mass + g.x + d.fa[0];
        dx[X_VEL + 1] = o1 * d.inv_mass + g.y + d.fa[1];
        dx[X_VEL + 2] = o2 * d.inv_mass + g.z + d.fa[2];
    }
    {
        float v[3];
#pragma unroll
        for(int i = 0; i < 3; i++){
            if constexpr(AXIAL) v[i] = p.c(C_JD + i) * (i == 0 ? w0 : (i == 1 ? w1 : w2));
            else v[i] = p.c(dyn_j_slot(C_JD, C_JOFF, i, 0)) * w0 + p.c(dyn_j_slot(C_JD, C_JOFF, i, 1)) * w1 + p.c(dyn_j_slot(C_JD, C_JOFF, i, 2)) * w2;
        }
        const float t0 = torque[0] - (w1 * v[2] - w2 * v[1]);
        const float t1 = torque[1] - (w2 * v[0] - w0 * v[2]);
        const float t2 = torque[2] - (w0 * v[1] - w1 * v[0]);
#pragma unroll
        for(int i = 0; i < 3; i++){
            if constexpr(AXIAL) dx[X_OMEGA + i] = p.c(C_JID + i) * (i == 0 ? t0 : (i == 1 ? t1 : t2)) + d.ta[i];
            else dx[X_OMEGA + i] = p.c(dyn_j_slot(C_JID, C_JIOFF, i, 0)) * t0 + p.c(dyn_j_slot(C_JID, C_JIOFF, i, 1)) * t1 + p.c(dyn_j_slot(C_JID, C_JIOFF, i, 2)) * t2 + d.ta[i];
        }
    }
    {   // first-order motor lag with separate rising / falling time constants: (setpoint - rpm) / tau(sign of the difference) = m d + h |d|
        const float4 m = p.c4(C_TAU_M), h = p.c4(C_TAU_H);
        const float mm[4] = {m.x, m.y, m.z, m.w}, hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for(int r = 0; r < 4; r++){
            const float dd = setpoint[r] - x[X_RPM + r];
            dx[X_RPM + r] = fmaf(hh[r], fabsf(dd), mm[r] * dd);
        }
    }
}

// ---- packed form of the axial vehicle's step: the integrated state of ONE environment as natural pairs -------------------------------------------
//   s[0] = (p0, p1)   s[1] = (v0, v1)   s[2] = (v2, w2)   s[3] = (q0, q1)   s[4] = (q2, q3)   s[5] = (w0, w1)   s[6] = (rpm0, rpm1)   s[7] = (rpm2, rpm3)
//   and p2 as a scalar.  Every pair is what one packed instruction of the dynamics produces or consumes (thrust curves and motor lag of two
//   rotors, torque rows 0 | 1, the x | y accelerations, J w and J^-1 (.) of axes 0 | 1), so the state algebra of RK4 (50_state_algebra.h) and
//   about half of a dynamics evaluation issue as FFMA2 / FADD2 / FMUL2: 55 instead of ~110 instructions per evaluation.
enum PackedSlot : int { K_P01 = 0, K_V01 = 1, K_VW2 = 2, K_Q01 = 3, K_Q23 = 4, K_W01 = 5, K_R01 = 6, K_R23 = 7, K_PAIRS = 8 };
template <class PC>
__device__ __forceinline__ void dynamics_axial_packed(const PC& p, const DynInvariants& d, const F2* __restrict__ s, F2 sp01, F2 sp23, F2* __restrict__ k, float& k_p2){
    using namespace p2;
    const F2 R01 = s[K_R01], R23 = s[K_R23];
    const float4 c0 = p.c4(C_COEF), c1 = p.c4(C_COEF + 4), c2 = p.c4(C_COEF + 8);
    const F2 tm01 = fma(fma(lo2(c2), R01, lo2(c1)), R01, lo2(c0));      // thrust of rotors 0 | 1 (60_dynamics.h:31: c0 + c1 rpm + c2 rpm^2)
    const F2 tm23 = fma(fma(hi2(c2), R23, hi2(c1)), R23, hi2(c0));
    const F2 tsum = add(tm01, tm23);
    const float T = tsum.x + tsum.y;
    const float4 a01 = p.c4(C_AT01), a23 = p.c4(C_AT01 + 4), a2 = p.c4(C_AT2);
    const F2 t01 = fma(hi2(a23), bc(tm23.y), fma(lo2(a23), bc(tm23.x), fma(hi2(a01), bc(tm01.y), mul(lo2(a01), bc(tm01.x)))));   // torque rows 0 | 1
    const F2 t2p = fma(hi2(a2), tm23, mul(lo2(a2), tm01));
    const float t2 = t2p.x + t2p.y;
    k[K_P01] = s[K_V01]; k_p2 = s[K_VW2].x;
    const float q0 = s[K_Q01].x, q1 = s[K_Q01].y, q2 = s[K_Q23].x, q3 = s[K_Q23].y;
    const float w0 = s[K_W01].x, w1 = s[K_W01].y, w2 = s[K_VW2].y;
    {
        const F2 h01 = mul(s[K_W01], bc(0.5f));       // exact scaling: the same values as (...) * 0.5
        const float h0 = h01.x, h1 = h01.y, h2 = w2 * 0.5f;
        k[K_Q01] = mk(fmaf(-q3, h2, fmaf(-q2, h1, -q1 * h0)), fmaf(-q3, h1, fmaf(q2, h2, q0 * h0)));
        k[K_Q23] = mk(fmaf(-q1, h2, fmaf(q3, h0, q0 * h1)), fmaf(-q2, h0, fmaf(q1, h1, q0 * h2)));
    }
    float dv2;
    {   // rotate (0, 0, T) by q, / m, + g + F_d / m
        const float T2 = T + T;
        const float v0 = q2 * T2, v1 = -q1 * T2;
        const float o0 = fmaf(v0, q0, -q3 * v1), o1 = fmaf(v1, q0, q3 * v0), o2 = fmaf(-q2, v0, q1 * v1) + T;
        k[K_V01] = fma(mk(o0, o1), bc(d.inv_mass), mk(d.ga[0], d.ga[1]));
        dv2 = fmaf(o2, d.inv_mass, d.ga[2]);
    }
    {   // J^-1 (tau - w x J w) + J^-1 tau_d
        const float4 jd = p.c4(C_JD), ji = p.c4(C_JID);
        const F2 v01 = mul(lo2(jd), s[K_W01]);
        const float v2 = jd.z * w2;
        const float e0 = fmaf(-w1, v2, fmaf(w2, v01.y, t01.x));
        const float e1 = fmaf(-w2, v01.x, fmaf(w0, v2, t01.y));
        const float e2 = fmaf(-w0, v01.y, fmaf(w1, v01.x, t2));
        k[K_W01] = fma(lo2(ji), mk(e0, e1), mk(d.ta[0], d.ta[1]));
        k[K_VW2] = mk(dv2, fmaf(ji.z, e2, d.ta[2]));
    }
    {
        const float4 m = p.c4(C_TAU_M), h = p.c4(C_TAU_H);
        const F2 d01 = sub(sp01, R01), d23 = sub(sp23, R23);
        k[K_R01] = fma(lo2(h), abs2(d01), mul(lo2(m), d01));
        k[K_R23] = fma(hi2(h), abs2(d23), mul(hi2(m), d23));
    }
}

// env_step twin for the compiled block (NOISE: action noise as in env_step; Langevin target as in env_step)
// FAST: default-math variant (min/max clamps, MUFU reciprocal square root for the quaternion, MUFU Box-Muller for the Langevin target)
// langevin_normals: the three N(0, 1) draws of the Langevin update, drawn by the caller ahead of time (legal only when nothing else draws from the
// stream in between, i.e. NOISE == false: the values depend on the stream alone, not on the action) or nullptr = draw here
// Langevin target of the trajectory (70_post_integration.h:127-170), the last part of post_integration: a function of its own so that a kernel can
// place it after the next step's observation has been handed to the tensor core (WITH_LANGEVIN = false in env_step_compiled, see k_rollout_raptor_ts)
template <class Spec, bool FAST, class PC>
__device__ __forceinline__ void langevin_update_compiled(EnvState<Spec>& st, const PC& p, uint64_t& rng, float dt, const float* __restrict__ langevin_normals = nullptr){
    if constexpr(Spec::LANGEVIN && FAST && B200L2F_LANGEVIN_BRANCH_FREE){
        // without a branch: every lane runs the update on copies, the results are committed by selects.  In a warp of mixed trajectory types the branchy
        // form executes the same instructions anyway, but as a separate basic block -- the 72 dependent integer operations of the six xorshift draws then
        // cannot be interleaved with the arithmetic around them.  A lane whose trajectory is not Langevin keeps its stream untouched.
        const bool lang = st.traj_type == 1;
        const float gamma = p[P_LANGEVIN_GAMMA], omega = p[P_LANGEVIN_OMEGA], sigma = p[P_LANGEVIN_SIGMA], alpha = p[P_LANGEVIN_ALPHA];
        const float sqrt_dt = p.c(C_SQRT_DT);
        uint64_t r2 = rng;
        float L[12];
#pragma unroll
        for(int i = 0; i < 12; i++) L[i] = st.lang[i];
#pragma unroll
        for(int dim = 0; dim < 3; dim++){
            const float x_prev = L[6 + dim], v_prev = L[9 + dim];
            const float dW = sqrt_dt * (langevin_normals ? langevin_normals[dim] : rng_normal_draw_fast(r2, 0.0f, 1.0f));
            const float v_next = v_prev + (-gamma * v_prev - omega * omega * x_prev) * dt + sigma * dW;
            const float x_next = x_prev + v_next * dt;
            L[6 + dim] = x_next; L[9 + dim] = v_next;
            const float v_smooth = alpha * v_next + (1.0f - alpha) * L[3 + dim];
            L[dim] = L[dim] + v_smooth * dt;
            L[3 + dim] = v_smooth;
        }
#pragma unroll
        for(int i = 0; i < 12; i++) st.lang[i] = lang ? L[i] : st.lang[i];
        rng = lang ? r2 : rng;
    }
    else if constexpr(Spec::LANGEVIN){
        if(st.traj_type == 1){
            const float gamma = p[P_LANGEVIN_GAMMA], omega = p[P_LANGEVIN_OMEGA], sigma = p[P_LANGEVIN_SIGMA], alpha = p[P_LANGEVIN_ALPHA];
            const float sqrt_dt = p.c(C_SQRT_DT);
#pragma unroll
            for(int dim = 0; dim < 3; dim++){
                const float x_prev = st.lang[6 + dim], v_prev = st.lang[9 + dim];
                const float dW = sqrt_dt * (langevin_normals ? langevin_normals[dim] : rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, 1.0f));
                const float v_next = v_prev + (-gamma * v_prev - omega * omega * x_prev) * dt + sigma * dW;
                const float x_next = x_prev + v_next * dt;
                st.lang[6 + dim] = x_next; st.lang[9 + dim] = v_next;
                const float v_smooth = alpha * v_next + (1.0f - alpha) * st.lang[3 + dim];
                st.lang[dim] = st.lang[dim] + v_smooth * dt;
                st.lang[3 + dim] = v_smooth;
            }
        }
    }
}

template <class Spec, bool ROLLED_RK4 = false, bool NOISE = false, bool FAST = false, bool AXIAL = false, bool WITH_LANGEVIN = true, class PC>
__device__ __forceinline__ void env_step_compiled(EnvState<Spec>& st, const PC& p, const DynInvariants& d, const float* __restrict__ action, uint64_t& rng,
                                                  float* __restrict__ hist_ptr, size_t n, const float* __restrict__ langevin_normals = nullptr){
    float setpoint[4];
    const float amin = p.c(C_G + 3), amax = p.c(C_JD + 3);
#pragma unroll
    for(int i = 0; i < 4; i++){
        float a = action[i];
        if constexpr(NOISE) a += rng_normal_t<Spec::RNG_OOL, FAST>(rng, 0.0f, p[P_ACTION_NOISE]);
        setpoint[i] = clamp_t<FAST>(a, -1.0f, 1.0f) * d.half_range + amin + d.half_range;
    }
    const float4 dts = p.c4(C_DT);
    const float dt = dts.x, dt2 = dts.y, dt3 = dts.z, dt6 = dts.w;
    constexpr bool PACKED_STEP = FAST && AXIAL && B200L2F_PACKED_FP32 && B200L2F_PACKED_DYNAMICS;
    if constexpr(PACKED_STEP){
        using namespace p2;
        F2 s[K_PAIRS], acc[K_PAIRS], tmp[K_PAIRS], k[K_PAIRS];
        s[K_P01] = mk(st.x[X_POS], st.x[X_POS + 1]); s[K_V01] = mk(st.x[X_VEL], st.x[X_VEL + 1]); s[K_VW2] = mk(st.x[X_VEL + 2], st.x[X_OMEGA + 2]);
        s[K_Q01] = mk(st.x[X_ORI], st.x[X_ORI + 1]); s[K_Q23] = mk(st.x[X_ORI + 2], st.x[X_ORI + 3]); s[K_W01] = mk(st.x[X_OMEGA], st.x[X_OMEGA + 1]);
        s[K_R01] = mk(st.x[X_RPM], st.x[X_RPM + 1]); s[K_R23] = mk(st.x[X_RPM + 2], st.x[X_RPM + 3]);
        const float p2_0 = st.x[X_POS + 2];
        float p2_acc, k_p2;
        const F2 sp01 = mk(setpoint[0], setpoint[1]), sp23 = mk(setpoint[2], setpoint[3]);
        const F2 c1 = bc(dt), c2 = bc(dt2), c3 = bc(dt3), c6 = bc(dt6);
        if constexpr(ROLLED_RK4){   // four stages as one loop body (compact-code kernels); same arithmetic and accumulation order
#pragma unroll
            for(int i = 0; i < K_PAIRS; i++){ acc[i] = s[i]; tmp[i] = s[i]; }
            p2_acc = p2_0;
#pragma unroll 1
            for(int stage = 0; stage < 4; stage++){
                dynamics_axial_packed(p, d, tmp, sp01, sp23, k, k_p2);
                const float wa = (stage == 0 || stage == 3) ? dt6 : dt3, wt = (stage == 2) ? dt : dt2;
                p2_acc = fmaf(wa, k_p2, p2_acc);
#pragma unroll
                for(int i = 0; i < K_PAIRS; i++){ acc[i] = fma(bc(wa), k[i], acc[i]); tmp[i] = fma(bc(wt), k[i], s[i]); }
            }
        }
        else{
            // the dynamics do not read the position: its stage inputs are never formed, only the accumulation (next = x + dt/6 k1 + dt/3 k2 + dt/3 k3 + dt/6 k4)
            dynamics_axial_packed(p, d, s, sp01, sp23, k, k_p2);
            p2_acc = fmaf(dt6, k_p2, p2_0); acc[0] = fma(c6, k[0], s[0]);
#pragma unroll
            for(int i = 1; i < K_PAIRS; i++){ acc[i] = fma(c6, k[i], s[i]); tmp[i] = fma(c2, k[i], s[i]); }
            dynamics_axial_packed(p, d, tmp, sp01, sp23, k, k_p2);
            p2_acc = fmaf(dt3, k_p2, p2_acc); acc[0] = fma(c3, k[0], acc[0]);
#pragma unroll
            for(int i = 1; i < K_PAIRS; i++){ acc[i] = fma(c3, k[i], acc[i]); tmp[i] = fma(c2, k[i], s[i]); }
            dynamics_axial_packed(p, d, tmp, sp01, sp23, k, k_p2);
            p2_acc = fmaf(dt3, k_p2, p2_acc); acc[0] = fma(c3, k[0], acc[0]);
#pragma unroll
            for(int i = 1; i < K_PAIRS; i++){ acc[i] = fma(c3, k[i], acc[i]); tmp[i] = fma(c1, k[i], s[i]); }
            dynamics_axial_packed(p, d, tmp, sp01, sp23, k, k_p2);
            p2_acc = fmaf(dt6, k_p2, p2_acc);
#pragma unroll
            for(int i = 0; i < K_PAIRS; i++) acc[i] = fma(c6, k[i], acc[i]);
        }
        {   // post integration (70_post_integration.h:20-37): unit quaternion, limits
            const F2 n2 = fma(acc[K_Q23], acc[K_Q23], mul(acc[K_Q01], acc[K_Q01]));
            const float inv = rsqrt_approx(n2.x + n2.y);
            acc[K_Q01] = mul(acc[K_Q01], bc(inv)); acc[K_Q23] = mul(acc[K_Q23], bc(inv));
        }
        st.x[X_POS] = acc[K_P01].x; st.x[X_POS + 1] = acc[K_P01].y; st.x[X_POS + 2] = p2_acc;
        st.x[X_ORI] = acc[K_Q01].x; st.x[X_ORI + 1] = acc[K_Q01].y; st.x[X_ORI + 2] = acc[K_Q23].x; st.x[X_ORI + 3] = acc[K_Q23].y;
        st.x[X_VEL] = acc[K_V01].x; st.x[X_VEL + 1] = acc[K_V01].y; st.x[X_VEL + 2] = acc[K_VW2].x;
        st.x[X_OMEGA] = acc[K_W01].x; st.x[X_OMEGA + 1] = acc[K_W01].y; st.x[X_OMEGA + 2] = acc[K_VW2].y;
        st.x[X_RPM] = acc[K_R01].x; st.x[X_RPM + 1] = acc[K_R01].y; st.x[X_RPM + 2] = acc[K_R23].x; st.x[X_RPM + 3] = acc[K_R23].y;
        // clamp of position / velocities to +-1e5: one range test (three-input maxima), the nine clamps only when it fails
        const float big = max3(max3(fabsf(st.x[0]), fabsf(st.x[1]), fabsf(st.x[2])), max3(fabsf(st.x[7]), fabsf(st.x[8]), fabsf(st.x[9])),
                               max3(fabsf(st.x[10]), fabsf(st.x[11]), fabsf(st.x[12])));
        if(!(big <= 100000.0f)){
#pragma unroll
            for(int i = 0; i < 3; i++){
                st.x[X_POS + i] = clamp_t<true>(st.x[X_POS + i], -100000.0f, 100000.0f);
                st.x[X_VEL + i] = clamp_t<true>(st.x[X_VEL + i], -100000.0f, 100000.0f);
                st.x[X_OMEGA + i] = clamp_t<true>(st.x[X_OMEGA + i], -100000.0f, 100000.0f);
            }
        }
    }
    else{
    float k[X_DIM], tmp[X_DIM], acc[X_DIM];
    if constexpr(ROLLED_RK4){   // four stages as one loop body (a quarter of the code, same arithmetic and accumulation order)
#pragma unroll
        for(int i = 0; i < X_DIM; i++){ acc[i] = st.x[i]; tmp[i] = st.x[i]; }
#pragma unroll 1
        for(int s = 0; s < 4; s++){
            dynamics_compiled<AXIAL>(p, d, tmp, setpoint, k);
            const float wa = (s == 0 || s == 3) ? dt6 : dt3;
            const float wt = (s == 2) ? dt : dt2;
#pragma unroll
            for(int i = 0; i < X_DIM; i++){ acc[i] += wa * k[i]; tmp[i] = st.x[i] + wt * k[i]; }
        }
#pragma unroll
        for(int i = 0; i < X_DIM; i++) st.x[i] = acc[i];
    }
    else if constexpr(FAST && B200L2F_PACKED_FP32){
    // state algebra (50_state_algebra.h axpys) on the packed fp32 pipe: one FFMA2 (fma.rn.f32x2) advances two state components, so the 119
    // FMAs of the four stages issue as 63 instructions.  Same products and sums as the scalar form (an FMA per component), same bits.
    constexpr int X2 = (X_DIM + 1) / 2;
    float2 x2[X2], k2[X2], t2[X2], a2[X2];
#pragma unroll
    for(int i = 0; i < X2; i++) x2[i] = make_float2(st.x[2 * i], 2 * i + 1 < X_DIM ? st.x[2 * i + 1] : 0.0f);
    k2[X2 - 1].y = 0.0f;
    float* kf = reinterpret_cast<float*>(k2); float* tf = reinterpret_cast<float*>(t2);
    const float2 c6 = make_float2(dt6, dt6), c3 = make_float2(dt3, dt3), c2 = make_float2(dt2, dt2), c1 = make_float2(dt, dt);
    dynamics_compiled<AXIAL>(p, d, st.x, setpoint, kf);
#pragma unroll
    for(int i = 0; i < X2; i++){ a2[i] = __ffma2_rn(c6, k2[i], x2[i]); t2[i] = __ffma2_rn(c2, k2[i], x2[i]); }
    dynamics_compiled<AXIAL>(p, d, tf, setpoint, kf);
#pragma unroll
    for(int i = 0; i < X2; i++){ a2[i] = __ffma2_rn(c3, k2[i], a2[i]); t2[i] = __ffma2_rn(c2, k2[i], x2[i]); }
    dynamics_compiled<AXIAL>(p, d, tf, setpoint, kf);
#pragma unroll
    for(int i = 0; i < X2; i++){ a2[i] = __ffma2_rn(c3, k2[i], a2[i]); t2[i] = __ffma2_rn(c1, k2[i], x2[i]); }
    dynamics_compiled<AXIAL>(p, d, tf, setpoint, kf);
#pragma unroll
    for(int i = 0; i < X2; i++) a2[i] = __ffma2_rn(c6, k2[i], a2[i]);
#pragma unroll
    for(int i = 0; i < X_DIM; i++) st.x[i] = (i & 1) ? a2[i >> 1].y : a2[i >> 1].x;
    }
    else{
    dynamics_compiled<AXIAL>(p, d, st.x, setpoint, k);
#pragma unroll
    for(int i = 0; i < X_DIM; i++){ acc[i] = st.x[i] + dt6 * k[i]; tmp[i] = st.x[i] + dt2 * k[i]; }
    dynamics_compiled<AXIAL>(p, d, tmp, setpoint, k);
#pragma unroll
    for(int i = 0; i < X_DIM; i++){ acc[i] += dt3 * k[i]; tmp[i] = st.x[i] + dt2 * k[i]; }
    dynamics_compiled<AXIAL>(p, d, tmp, setpoint, k);
#pragma unroll
    for(int i = 0; i < X_DIM; i++){ acc[i] += dt3 * k[i]; tmp[i] = st.x[i] + dt * k[i]; }
    dynamics_compiled<AXIAL>(p, d, tmp, setpoint, k);
#pragma unroll
    for(int i = 0; i < X_DIM; i++) st.x[i] = acc[i] + dt6 * k[i];
    }
    {
        float nrm = 0.0f;
#pragma unroll
        for(int i = 0; i < 4; i++) nrm += st.x[X_ORI + i] * st.x[X_ORI + i];
        if constexpr(FAST){
            const float inv = rsqrt_approx(nrm);
#pragma unroll
            for(int i = 0; i < 4; i++) st.x[X_ORI + i] = st.x[X_ORI + i] * inv;
        }
        else{
            nrm = sqrtf(nrm);
#pragma unroll
            for(int i = 0; i < 4; i++) st.x[X_ORI + i] = st.x[X_ORI + i] / nrm;
        }
#pragma unroll
        for(int i = 0; i < 3; i++){
            st.x[X_POS + i] = clamp_t<FAST>(st.x[X_POS + i], -100000.0f, 100000.0f);
            st.x[X_VEL + i] = clamp_t<FAST>(st.x[X_VEL + i], -100000.0f, 100000.0f);
            st.x[X_OMEGA + i] = clamp_t<FAST>(st.x[X_OMEGA + i], -100000.0f, 100000.0f);
        }
    }
    }
#pragma unroll
    for(int i = 0; i < 4; i++) st.last_action[i] = action[i];
#pragma unroll
    for(int i = 0; i < 4; i++) st.x[X_RPM + i] = clamp_t<FAST>(st.x[X_RPM + i], amin, amax);
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) st.hist[i] = action[i];
    }
    else{
        const int cs = st.current_step;
#pragma unroll
        for(int i = 0; i < 4; i++) hist_ptr[(size_t)(4 * cs + i) * n] = action[i];
        st.current_step = (cs + 1) % Spec::H;
    }
    if constexpr(WITH_LANGEVIN) langevin_update_compiled<Spec, FAST>(st, p, rng, dt, langevin_normals);
}

// ---- shared-memory plan (bytes) ---------------------------------------------------------------------------------------
struct TcSmem {
    static constexpr int CHUNK = BLOCK * 16;                  // one K chunk (4 tf32 columns) of all 128 rows: 2 KB
    static constexpr int A_CHUNKS_HI = 12, A_CHUNKS_LO = 10;  // chunk pairs: (0,1)(2,3)(4,5) obs | x1 ; (6,7)(8,9) h ; (10,11) constants (hi plane only)
    static constexpr int A_HI = 0;
    static constexpr int A_LO = A_HI + A_CHUNKS_HI * CHUNK;
    static constexpr int B = A_LO + A_CHUNKS_LO * CHUNK;     // weight image (TMA destination, 16-byte aligned)
    static constexpr int DYN = B + TcImage::BYTES;
    static constexpr int BAR = DYN + C_DIM * BLOCK * 4;       // two mbarriers + TMEM base address
    static constexpr int TOTAL = BAR + 32;
};
static_assert(BLOCK == 128, "the tensor-core rollout maps one CTA to one M = 128 MMA tile");

// G1_TC: dense 1 on the tensor cores as well (two MMA round trips per step) or on the CUDA cores (one round trip: only the GRU GEMM,
// 1536 of the 1952 MACs, uses tcgen05).  Measured on B200 (profiles/): see DESIGN.md section 7.
template <class Spec, bool FAST, bool UNIFORM, bool G1_TC>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_rollout_raptor_tc(const __grid_constant__ RolloutArgs a, const float* __restrict__ tc_image){
    constexpr int HD = 16;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* a_hi = reinterpret_cast<float*>(smraw + TcSmem::A_HI);
    float* a_lo = reinterpret_cast<float*>(smraw + TcSmem::A_LO);
    float* sm_b = reinterpret_cast<float*>(smraw + TcSmem::B);
    float* sm_dyn = reinterpret_cast<float*>(smraw + TcSmem::DYN);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smraw + TcSmem::BAR);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x, warp = tc::uniform_warp_index();

    if(tid == 0){
        tc::mbar_init(bar_tma, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
    }
    if(warp == 0) tc::tmem_alloc<128>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if(tid == 0){
        tc::mbar_expect_tx(bar_tma, TcImage::BYTES);
        tc::tma_load_1d(sm_b, tc_image, TcImage::BYTES, bar_tma);
    }
    const int e = blockIdx.x * BLOCK + tid;
    const bool active = e < a.n;
    const size_t n = (size_t)a.n;
    const size_t env = active ? (size_t)e : 0;
    const ParamsCompiledT<UNIFORM> p = stage_dynamics_compiled<UNIFORM>(sm_dyn, a.params, n, env, a.row0);
    EnvState<Spec> st;
    load_state(st, a.state + env, n);
    DynInvariants d;
    {
        ParamsGlobal pg{a.params + env, n};
        dyn_invariants(d, pg, st);
    }
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = a.rng[env];
    float h[HD];
#pragma unroll
    for(int j = 0; j < HD; j++) h[j] = a.hidden[(size_t)j * n + env];
    int gs = a.gru_step[env];
    float ret = 0.0f; int eplen = 0; bool done = false;
    const bool no_auto_reset = a.no_auto_reset != 0;
    // this thread's row inside a K chunk: 16 bytes at chunk * 2048 + tid * 16  (8-row core matrices are 128 contiguous bytes: SBO = 128)
    float4* row_hi = reinterpret_cast<float4*>(a_hi) + tid;
    float4* row_lo = reinterpret_cast<float4*>(a_lo) + tid;
    constexpr int CH = BLOCK;   // float4 stride between chunks
    auto put4 = [&](int chunk, float v0, float v1, float v2, float v3){
        float h0, h1, h2, h3, l0, l1, l2, l3;
        tc::split_tf32(v0, h0, l0); tc::split_tf32(v1, h1, l1); tc::split_tf32(v2, h2, l2); tc::split_tf32(v3, h3, l3);
        row_hi[chunk * CH] = make_float4(h0, h1, h2, h3);
        row_lo[chunk * CH] = make_float4(l0, l1, l2, l3);
    };
    row_hi[10 * CH] = make_float4(1.0f, 1.0f, 0.0f, 0.0f);   // constant columns: two ones (b_hh / b_ih; G3 uses the first for b2), then zeros
    row_hi[11 * CH] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    put4(6, h[0], h[1], h[2], h[3]); put4(7, h[4], h[5], h[6], h[7]); put4(8, h[8], h[9], h[10], h[11]); put4(9, h[12], h[13], h[14], h[15]);

    // descriptors (issuing thread only uses them)
    const uint32_t a_hi_s = tc::smem_u32(a_hi), a_lo_s = tc::smem_u32(a_lo), b_s = tc::smem_u32(sm_b);
    constexpr uint32_t A_LBO = TcSmem::CHUNK, SBO = 128;
    constexpr uint32_t IDESC16 = tc::make_idesc_tf32(128, 16), IDESC64 = tc::make_idesc_tf32(128, 64);
    const uint32_t d1 = tmem_base + 0, d2 = tmem_base + 16, d3 = tmem_base + 80;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    uint32_t phase = 0;
    // one GEMM = `pairs` K=8 steps over the chunk pairs first_pair.. of A against consecutive chunk pairs of B (3 products per step)
    auto issue_gemm = [&](uint32_t dcol, int a_first_chunk_0, int n_pairs_0, int a_first_chunk_1, int n_pairs_1, int b_hi_off, int b_lo_off, uint32_t N, uint32_t idesc){
        uint32_t acc = 0;
        int bpair = 0;
        for(int seg = 0; seg < 2; seg++){
            const int first = seg == 0 ? a_first_chunk_0 : a_first_chunk_1;
            const int cnt = seg == 0 ? n_pairs_0 : n_pairs_1;
            for(int i = 0; i < cnt; i++, bpair++){
                const int chunk = first + 2 * i;
                const uint64_t ahi = tc::make_smem_desc(a_hi_s + chunk * TcSmem::CHUNK, A_LBO, SBO);
                const uint64_t bhi = tc::make_smem_desc(b_s + b_hi_off * 4 + bpair * 2 * N * 16, N * 16, SBO);
                const uint64_t blo = tc::make_smem_desc(b_s + b_lo_off * 4 + bpair * 2 * N * 16, N * 16, SBO);
                tc::mma_tf32(dcol, ahi, bhi, idesc, acc); acc = 1;
                tc::mma_tf32(dcol, ahi, blo, idesc, 1);
                if(chunk < TcSmem::A_CHUNKS_LO){   // the constant chunks have no lo plane (their lo part is exactly zero)
                    const uint64_t alo = tc::make_smem_desc(a_lo_s + chunk * TcSmem::CHUNK, A_LBO, SBO);
                    tc::mma_tf32(dcol, alo, bhi, idesc, 1);
                }
            }
        }
        tc::mma_commit(bar_mma);
    };
    tc::mbar_wait(bar_tma, 0);   // weight image landed (async proxy write -> visible to the tensor core, no generic-proxy reads needed except h0)
    __syncthreads();

    for(int t = 0; t < a.T; t++){
        if(a.out_states && active && (t % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(t / a.state_stride) * n + env) * Spec::STATE_DIM);
        float obs[22];
        observe18<Spec, false>(st, p, rng, obs);
        if constexpr(Spec::H == 1){
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = st.hist[i];
        }
        else{
            const int cur = st.current_step == 0 ? Spec::H - 1 : st.current_step - 1;
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = hist_ptr[(size_t)(4 * cur + i) * n];
        }
        if(a.out_obs && active){
            float* row = a.out_obs + ((size_t)t * n + env) * 22;
#pragma unroll
            for(int i = 0; i < 22; i++) row[i] = obs[i];
        }
        float x1[HD];
        if constexpr(G1_TC){
            // ---- G1: dense 1 on the tensor cores
            put4(0, obs[0], obs[1], obs[2], obs[3]); put4(1, obs[4], obs[5], obs[6], obs[7]); put4(2, obs[8], obs[9], obs[10], obs[11]);
            put4(3, obs[12], obs[13], obs[14], obs[15]); put4(4, obs[16], obs[17], obs[18], obs[19]); put4(5, obs[20], obs[21], 1.0f, 0.0f);
            tc::fence_async_smem();
            tc::tc_fence_before();
            __syncthreads();
            if(tid == 0){
                tc::tc_fence_after();
                issue_gemm(d1, 0, 3, 0, 0, TcImage::B1_HI, TcImage::B1_LO, 16, IDESC16);
            }
            tc::mbar_wait(bar_mma, phase); phase ^= 1;
            tc::tc_fence_after();
            tc::tmem_ld16(d1 + lane_off, x1);
            tc::tmem_ld_wait();
        }
        else{
            // ---- dense 1 (22 -> 16, 352 MACs) on the CUDA cores: one tensor-core round trip less per step
            const float* w1 = sm_b + TcImage::W1T;
#pragma unroll
            for(int j = 0; j < HD; j++) x1[j] = w1[22 * HD + j];
#pragma unroll
            for(int k = 0; k < 22; k++){
#pragma unroll
                for(int j4 = 0; j4 < HD / 4; j4++){
                    const float4 w = *reinterpret_cast<const float4*>(w1 + k * HD + 4 * j4);
                    x1[4 * j4] += w.x * obs[k]; x1[4 * j4 + 1] += w.y * obs[k]; x1[4 * j4 + 2] += w.z * obs[k]; x1[4 * j4 + 3] += w.w * obs[k];
                }
            }
        }
#pragma unroll
        for(int j = 0; j < HD; j++) x1[j] = fmaxf(x1[j], 0.0f);
        // reset_truncate (gru/operations_generic.h:76-86): the hidden rows of A are rewritten when the counter wrapped
        if(!no_auto_reset && gs >= a.seq_len){
#pragma unroll
            for(int j = 0; j < HD; j++) h[j] = sm_b[TcImage::H0 + j];
            put4(6, h[0], h[1], h[2], h[3]); put4(7, h[4], h[5], h[6], h[7]); put4(8, h[8], h[9], h[10], h[11]); put4(9, h[12], h[13], h[14], h[15]);
            gs = 0;
        }
        // ---- G2: GRU pre-activations
        put4(0, x1[0], x1[1], x1[2], x1[3]); put4(1, x1[4], x1[5], x1[6], x1[7]); put4(2, x1[8], x1[9], x1[10], x1[11]); put4(3, x1[12], x1[13], x1[14], x1[15]);
        tc::fence_async_smem();
        tc::tc_fence_before();
        __syncthreads();
        if(tid == 0){
            tc::tc_fence_after();
            issue_gemm(d2, 0, 2, 6, 3, TcImage::B2_HI, TcImage::B2_LO, 64, IDESC64);
        }
        tc::mbar_wait(bar_mma, phase); phase ^= 1;
        tc::tc_fence_after();
        float hn[HD];
        {
            float r[HD], nx[HD], nh[HD];
            tc::tmem_ld16(d2 + lane_off + 0, r);
            tc::tmem_ld16(d2 + lane_off + 32, nx);
            tc::tmem_ld16(d2 + lane_off + 48, nh);
            tc::tmem_ld_wait();
#pragma unroll
            for(int j = 0; j < HD; j++) nx[j] = tanhf_<FAST>(nx[j] + nh[j] * sigmoidf_<FAST>(r[j]));
            float z[HD];
            tc::tmem_ld16(d2 + lane_off + 16, z);
            tc::tmem_ld_wait();
#pragma unroll
            for(int j = 0; j < HD; j++){ const float zz = sigmoidf_<FAST>(z[j]); hn[j] = (1.0f - zz) * nx[j] + zz * h[j]; }
        }
        // ---- dense 2 (16 -> 4, 64 MACs): on the CUDA cores -- a third tensor-core round trip per step costs more latency than 64 FFMAs
        float act[4];
        {
            const float* w2 = sm_b + TcImage::W2T;
            const float4 b = *reinterpret_cast<const float4*>(w2 + 64);
            act[0] = b.x; act[1] = b.y; act[2] = b.z; act[3] = b.w;
#pragma unroll
            for(int k = 0; k < HD; k++){
                const float4 w = *reinterpret_cast<const float4*>(w2 + 4 * k);
                act[0] += w.x * hn[k]; act[1] += w.y * hn[k]; act[2] += w.z * hn[k]; act[3] += w.w * hn[k];
            }
        }
        {   // gru/operations_generic.h:400-410: this step's output is kept, the stored state resets when the counter wraps
            const int new_step = gs + 1;
            const bool wrap = !no_auto_reset && new_step >= a.seq_len;
#pragma unroll
            for(int j = 0; j < HD; j++) h[j] = wrap ? sm_b[TcImage::H0 + j] : hn[j];
            gs = wrap ? 0 : new_step;
            // the hidden rows of the A operand for the next step's G2 (its MMAs are issued only after the next barrier)
            put4(6, h[0], h[1], h[2], h[3]); put4(7, h[4], h[5], h[6], h[7]); put4(8, h[8], h[9], h[10], h[11]); put4(9, h[12], h[13], h[14], h[15]);
        }
        if(a.out_actions && active) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env) * 4) = make_float4(act[0], act[1], act[2], act[3]);
        RewardInputs ri;
        reward_inputs(ri, st);
        if(Spec::H == 1 || active) env_step_compiled<Spec>(st, p, d, act, rng, hist_ptr, n);
        const bool term = env_terminated(p, st.x);
        const float rw = env_reward(p, ri, act, st.x, term, d.dt);
        if(a.out_rewards && active) a.out_rewards[(size_t)t * n + env] = rw;
        if(a.out_term && active) a.out_term[(size_t)t * n + env] = term ? 1 : 0;
        if(!done){ ret += rw; eplen += 1; done = term; }
    }
    if(active){
        if(a.out_states && (a.T % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(a.T / a.state_stride) * n + env) * Spec::STATE_DIM);
        store_state(st, a.state + env, n);
        a.rng[env] = rng;
#pragma unroll
        for(int j = 0; j < HD; j++) a.hidden[(size_t)j * n + env] = h[j];
        a.gru_step[env] = gs;
        if(a.out_returns) a.out_returns[env] = ret;
        if(a.out_eplen) a.out_eplen[env] = eplen;
        if(a.out_done) a.out_done[env] = done ? 1 : 0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if(warp == 0) tc::tmem_dealloc<128>(tmem_base);
}


// =================================================================================================================================
// TMEM-A variant ("TS" MMAs): the A operand lives in tensor memory as well.  Every thread writes ITS row of the activations (hi and lo
// planes) with tcgen05.st into its TMEM lane, the GRU hidden state lives ONLY in TMEM (hi + lo == h exactly in fp32; the tensor core
// ignores the 13 low mantissa bits it cannot use), and shared memory holds just the weight image and the staged dynamics block:
// 63 KB per CTA instead of 107 KB, 128 TMEM columns per CTA -> THREE CTAs (12 warps) per SM instead of two.
//   TMEM columns (128):  [0,16) h hi | [16,32) h lo | [32,56) obs hi | [56,80) obs lo | [80,96) D1 | later: [32,48) x1 hi | [48,64) x1 lo | [64,128) D2
//   lifetimes make the overlaps safe: obs/D1 are dead before x1 is written, x1/h are read by G2 while it writes D2 into [64,128).
// G1 folds its bias in as K column 22 (= 1.0); G2 has no spare K column (K = 32 exactly), its biases are added in the epilogue.
// =================================================================================================================================
template <bool AXIAL, int T_CTAS = 3>
struct TsSmemT {
    static_assert(T_CTAS == 3 || (T_CTAS == 4 && AXIAL), "four CTAs per SM need the compact (axial) dynamics block");
    static constexpr int CTAS = T_CTAS;                                       // resident CTAs per SM this variant is built for (3: 168 registers, 4: 128 registers)
    static constexpr int DSTRIDE = CTAS == 4 ? C_DIM_AXIAL : C_DIM;           // 4 CTAs/SM: 28 KB image + 22.5 KB compact block = 50.6 KB per CTA
    static constexpr int B = 0;                                // weight image (TMA destination)
    static constexpr int DYN = B + TcImage::BYTES;
    static constexpr int BAR = DYN + DSTRIDE * BLOCK * 4;
    static constexpr int HID = BAR + 32;                       // B200L2F_H_IN_SMEM: h as float4 columns [4][BLOCK] (conflict-free LDS.128 / STS.128)
    static constexpr int TOTAL = HID + (B200L2F_H_IN_SMEM ? 16 * BLOCK * 4 : 0);
};
// NOISE: observation / action noise (18 + 4 normal draws per step, each skipped when its std is 0) with the MUFU Box-Muller.
// CTAS: resident CTAs per SM the instantiation is register-budgeted for.  3 (168 registers) wins while the tiles do not fill 4 x SMs slots
// (65 536 environments = 512 tiles: 16.1e9 vs 14.8e9 env-steps/s); 4 (128 registers, compact dynamics block, 72 B of spills) wins once they do
// (1 048 576 environments: 18.3e9 vs 16.5e9) -- profiles/r02_exp2_*.log.  The launcher picks by tile count.
// RECORD = false: the launch asks for no per-step output (states / observations / actions / rewards / terminated flags; returns and episode lengths are
// per-item outputs and always available): the five uniform `if(a.out_*)` tests of a step disappear, and with them five basic-block boundaries the
// instruction scheduler could not move code across.
template <class Spec, bool FAST, bool UNIFORM, bool AXIAL, bool NOISE = false, int CTAS = 3, bool RECORD = true>
__global__ void __launch_bounds__(BLOCK, CTAS) k_rollout_raptor_ts(const __grid_constant__ RolloutArgs a, const float* __restrict__ tc_image){
    using TsSmem = TsSmemT<AXIAL, CTAS>;
    static_assert(FAST, "the TMEM-A kernel reads the scaled-gate image (build_tc_image_host(..., true)): default math only");
    constexpr int HD = 16;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* sm_b = reinterpret_cast<float*>(smraw + TsSmem::B);
    float* sm_dyn = reinterpret_cast<float*>(smraw + TsSmem::DYN);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smraw + TsSmem::BAR);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x, warp = tc::uniform_warp_index();
    if(tid == 0){
        tc::mbar_init(bar_tma, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
    }
    if(warp == 0) tc::tmem_alloc<128>(tmem_slot);
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
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    __shared__ int s_item;
    constexpr uint32_t C_H_HI = 0, C_H_LO = 16, C_OBS_HI = 32, C_OBS_LO = 56, C_D1 = 80, C_X1_HI = 32, C_X1_LO = 48, C_D2 = 64;
    // write 8 values as a hi block and a lo block (8 columns each) of this thread's lane
    auto put8 = [&](uint32_t col_hi, uint32_t col_lo, const float* v){
        float hi[8], lo[8];
        if constexpr(B200L2F_SPLIT_PACKED){
            const float2 minus1 = make_float2(-1.0f, -1.0f);
#pragma unroll
            for(int i = 0; i < 8; i += 2){
                hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u); hi[i + 1] = __uint_as_float(__float_as_uint(v[i + 1]) & 0xFFFFE000u);
                const float2 l = __ffma2_rn(make_float2(hi[i], hi[i + 1]), minus1, make_float2(v[i], v[i + 1]));   // x - hi, exact
                lo[i] = l.x; lo[i + 1] = l.y;
            }
        }
        else{
#pragma unroll
            for(int i = 0; i < 8; i++) tc::split_tf32(v[i], hi[i], lo[i]);
        }
        tc::tmem_st8(tmem_base + lane_off + col_hi, hi);
        tc::tmem_st8(tmem_base + lane_off + col_lo, lo);
    };
    // same for values that already sit in register pairs (TMEM loads, packed gate results): the low parts as x - hi on the packed pipe (exact)
    auto put8p = [&](uint32_t col_hi, uint32_t col_lo, const float* v){
        if constexpr(B200L2F_SPLIT_PAIRS){
            float hi[8], lo[8];
#pragma unroll
            for(int i = 0; i < 8; i += 2){
                hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u); hi[i + 1] = __uint_as_float(__float_as_uint(v[i + 1]) & 0xFFFFE000u);
                const float2 l = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(-hi[i], -hi[i + 1]));
                lo[i] = l.x; lo[i + 1] = l.y;
            }
            tc::tmem_st8(tmem_base + lane_off + col_hi, hi);
            tc::tmem_st8(tmem_base + lane_off + col_lo, lo);
        }
        else put8(col_hi, col_lo, v);
    };
    float4* sm_h = reinterpret_cast<float4*>(smraw + TsSmem::HID) + tid;   // this thread's hidden state: sm_h[q * BLOCK], q = 0..3
    auto keep_h = [&](const float* v){
        if constexpr(B200L2F_H_IN_SMEM){
#pragma unroll
            for(int q = 0; q < 4; q++) sm_h[q * BLOCK] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    };
    auto load_h = [&](float* v){
#pragma unroll
        for(int q = 0; q < 4; q++){ const float4 f = sm_h[q * BLOCK]; v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w; }
    };
    const uint32_t b_s = tc::smem_u32(sm_b);
    constexpr uint32_t SBO = 128;
    constexpr uint32_t IDESC16 = tc::make_idesc_tf32(128, 16), IDESC64 = tc::make_idesc_tf32(128, 64);
    uint32_t phase = 0;
    // ksteps K=8 steps: A columns [a_hi + 8 s, +8) / [a_lo + 8 s, +8) against B chunk pairs (b_pair0 + s)
    // the B descriptors differ only in their start address: two bases (N = 16 / N = 64 tiles), every other one is a compile-time offset away
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
    tc::mbar_wait(bar_tma, 0);
    __syncthreads();

    // ---- persistent work loop over (tile, time-chunk) items.  Items are handed out in chunk-major order by one atomic counter; an item may
    // ---- start once the previous chunk of the same tile has been published (that item has a smaller index, i.e. it was claimed earlier by
    // ---- a CTA that is already running, so waiting can never deadlock).  With one chunk per tile this degenerates to a plain tile loop.
    const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
    const int n_chunks = a.n_chunks;
    const int total_items = n_tiles * n_chunks;
    for(;;){
    if(tid == 0) s_item = atomicAdd(a.sched, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if(item >= total_items) break;
    const int tile = item % n_tiles, chunk = item / n_tiles;
    if(chunk > 0){
        if(tid == 0){
            const int* prog = a.sched + 1 + tile;
            int v;
            do{ asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog) : "memory"); if(v < chunk) __nanosleep(64); } while(v < chunk);
        }
        __syncthreads();
        __threadfence();   // gpu-scope fence in every thread: drops stale L1 lines of the tile's buffers (another SM wrote them during this launch)
    }
    const int t_begin = chunk * a.chunk_steps;
    const int t_end = min(a.T, t_begin + a.chunk_steps);
    const int e = tile * BLOCK + tid;
    const bool active = e < a.n;
    const size_t env = active ? (size_t)e : 0;
    const ParamsCompiledT<UNIFORM> p = stage_dynamics_compiled<UNIFORM, true, false, TsSmem::DSTRIDE>(sm_dyn, a.params, n, env, a.row0);
    EnvState<Spec> st;
    load_state_cg(st, a.state + env, n);
    DynInvariants d;
    {
        ParamsGlobal pg{a.params + env, n};
        dyn_invariants(d, pg, st);
    }
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = __ldcg(a.rng + env);
    int gs = __ldcg(a.gru_step + env);
    float ret = 0.0f; int eplen = 0; bool done = false;
    if(chunk > 0){ ret = __ldcg(a.acc_ret + env); const int v = __ldcg(a.acc_len + env); eplen = v >> 1; done = (v & 1) != 0; }
    {
        float h[HD];
#pragma unroll
        for(int j = 0; j < HD; j++) h[j] = __ldcg(a.hidden + (size_t)j * n + env);
        put8(C_H_HI, C_H_LO, h); put8(C_H_HI + 8, C_H_LO + 8, h + 8);
        keep_h(h);
    }
    tc::tmem_st_wait();

    // (Experiment, off by default: measured 16.27e9 against 16.97e9 env-steps/s -- with three CTAs per SM the round trip is already hidden by the other
    // CTAs, and the longer live ranges of the reward inputs across the barrier cost more than the shorter dependency chain gains.)
    // Software pipeline over the two tensor-core round trips of a step: once the integrator has produced the next state, the NEXT step's observation is
    // split, stored and handed to the tensor core (G1) BEFORE this step's termination test, reward and Langevin target update run -- those ~250
    // instructions then execute in the shadow of the G1 round trip instead of after it.  Legal when nothing between the two points feeds the
    // observation or the RNG order: no observation noise (its draws would have to follow the Langevin draws) and an observation that does not read the
    // trajectory target (RAPTOR layout; DEFAULT has no Langevin target).
    constexpr bool PIPELINE_G1 = B200L2F_PIPELINE_G1 && !NOISE && (Spec::OBS_LAYOUT == OBS_RAPTOR || !Spec::LANGEVIN);
    float x1_cuda[B200L2F_G1_CUDA != 0 ? HD : 1];
    auto start_g1 = [&](int t){      // observation of step t (state row / observation row recorded as of step t) -> TMEM -> dense 1 issued
        if(RECORD && a.out_states && active && (t % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(t / a.state_stride) * n + env) * Spec::STATE_DIM);
        float obs[24];
        observe18<Spec, NOISE, true>(st, p, rng, obs);
        if constexpr(Spec::H == 1){
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = st.hist[i];
        }
        else{
            const int cur = st.current_step == 0 ? Spec::H - 1 : st.current_step - 1;
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = hist_ptr[(size_t)(4 * cur + i) * n];
        }
        if(RECORD && a.out_obs && active){
            float* row = a.out_obs + ((size_t)t * n + env) * 22;
#pragma unroll
            for(int i = 0; i < 22; i++) row[i] = obs[i];
        }
        if constexpr(B200L2F_G1_CUDA != 0){
            // experiment (profiles/r02_exp11_dense1_cuda_cores.log): dense 1 (22 -> 16) on the packed fp32 pipe from the image's fp32 k-major copy of W1 -- 176 FFMA2 +
            // 88 broadcast LDS.128 instead of the operand split, a CTA barrier and a tensor-core round trip
            const float* w1 = sm_b + TcImage::W1T;
            float2 acc[8];
#pragma unroll
            for(int j = 0; j < 8; j++) acc[j] = make_float2(w1[22 * 16 + 2 * j], w1[22 * 16 + 2 * j + 1]);
#pragma unroll
            for(int k = 0; k < 22; k++){
                const float2 x = make_float2(obs[k], obs[k]);
#pragma unroll
                for(int q = 0; q < 4; q++){
                    const float4 w = *reinterpret_cast<const float4*>(w1 + 16 * k + 4 * q);
                    acc[2 * q] = __ffma2_rn(x, make_float2(w.x, w.y), acc[2 * q]);
                    acc[2 * q + 1] = __ffma2_rn(x, make_float2(w.z, w.w), acc[2 * q + 1]);
                }
            }
#pragma unroll
            for(int j = 0; j < 8; j++){ x1_cuda[2 * j] = acc[j].x; x1_cuda[2 * j + 1] = acc[j].y; }
            return;
        }
        obs[22] = 1.0f; obs[23] = 0.0f;   // bias column, pad
        // ---- G1: dense 1 (A = obs in TMEM)
        put8(C_OBS_HI, C_OBS_LO, obs); put8(C_OBS_HI + 8, C_OBS_LO + 8, obs + 8); put8(C_OBS_HI + 16, C_OBS_LO + 16, obs + 16);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        if(warp == 0 && tc::elect_one()){
            tc::tc_fence_after();
            issue_gemm(C_D1, C_OBS_HI, C_OBS_LO, 3, TcImage::B1_HI, TcImage::B1_LO, 0, 16, IDESC16, 0);
            tc::mma_commit(bar_mma);
        }
    };
    if(PIPELINE_G1 && t_begin < t_end) start_g1(t_begin);
    for(int t = t_begin; t < t_end; t++){
        if constexpr(!PIPELINE_G1) start_g1(t);
        // independent work in the shadow of the MMA round trip: the Langevin target's three normals depend on the RNG stream only
        constexpr bool HOIST = B200L2F_HOIST_LANGEVIN && Spec::LANGEVIN && !NOISE && !PIPELINE_G1;
        float lang_normals[3];
        if constexpr(HOIST){
            if(st.traj_type == 1){
#pragma unroll
                for(int dim = 0; dim < 3; dim++) lang_normals[dim] = rng_normal_t<Spec::RNG_OOL, true>(rng, 0.0f, 1.0f);
            }
        }
        float x1[HD];
        if constexpr(B200L2F_G1_CUDA != 0){
#pragma unroll
            for(int j = 0; j < HD; j++) x1[j] = x1_cuda[j];
        }
        else{
            tc::mbar_wait(bar_mma, phase); phase ^= 1;
            tc::tc_fence_after();
            tc::tmem_ld16(tmem_base + lane_off + C_D1, x1);
            tc::tmem_ld_wait();
        }
#pragma unroll
        for(int j = 0; j < HD; j += 2){                       // 2 ReLU(x1) = x1 + |x1|: one FADD2 per two units (the image's W_ih columns carry the 0.5: bit-identical)
            const float2 v = make_float2(x1[j], x1[j + 1]);
            const float2 r = __fadd2_rn(v, make_float2(fabsf(v.x), fabsf(v.y)));
            x1[j] = r.x; x1[j + 1] = r.y;
        }
        if(!no_auto_reset && gs >= a.seq_len){   // reset_truncate (gru/operations_generic.h:76-86)
            put8(C_H_HI, C_H_LO, sm_b + TcImage::H0); put8(C_H_HI + 8, C_H_LO + 8, sm_b + TcImage::H0 + 8);
            keep_h(sm_b + TcImage::H0);
            gs = 0;
        }
        // ---- G2: GRU pre-activations (A = [x1 | h] in TMEM, K = 32)
        put8p(C_X1_HI, C_X1_LO, x1); put8p(C_X1_HI + 8, C_X1_LO + 8, x1 + 8);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncthreads();
        if(warp == 0 && tc::elect_one()){
            tc::tc_fence_after();
            issue_gemm(C_D2, C_X1_HI, C_X1_LO, 2, TcImage::B2_HI, TcImage::B2_LO, 0, 64, IDESC64, 0);
            issue_gemm(C_D2, C_H_HI, C_H_LO, 2, TcImage::B2_HI, TcImage::B2_LO, 2, 64, IDESC64, 1);
            tc::mma_commit(bar_mma);
        }
        tc::mbar_wait(bar_mma, phase); phase ^= 1;
        tc::tc_fence_after();
        float hn[HD];
        {
            const float* bias = sm_b + TcImage::BIAS2;
            float r[HD], nx[HD], nh[HD];
            tc::tmem_ld16(tmem_base + lane_off + C_D2 + 0, r);
            tc::tmem_ld16(tmem_base + lane_off + C_D2 + 32, nx);
            tc::tmem_ld16(tmem_base + lane_off + C_D2 + 48, nh);
            tc::tmem_ld_wait();
            float z[HD], hh[HD], hl[HD];
            if constexpr(B200L2F_TMEM_PREFETCH){   // in flight while the MUFU chain below runs
                tc::tmem_ld16(tmem_base + lane_off + C_D2 + 16, z);
                if constexpr(!B200L2F_H_IN_SMEM){
                    tc::tmem_ld16(tmem_base + lane_off + C_H_HI, hh);
                    tc::tmem_ld16(tmem_base + lane_off + C_H_LO, hl);
                }
            }
            if constexpr(B200L2F_GATES_PACKED){
                // two hidden units per packed instruction; each lane of FADD2 / FFMA2 rounds like the scalar instruction, so the bits are those of
                // the scalar form below
                const float2 one2 = make_float2(1.0f, 1.0f), minus2 = make_float2(-2.0f, -2.0f);
#pragma unroll
                for(int j = 0; j < HD; j += 2){
                    const float2 b_r = *reinterpret_cast<const float2*>(bias + j), b_x = *reinterpret_cast<const float2*>(bias + 32 + j), b_h = *reinterpret_cast<const float2*>(bias + 48 + j);
                    const float2 tr = __fadd2_rn(make_float2(r[j], r[j + 1]), b_r);
                    const float2 dr = __fadd2_rn(one2, make_float2(ex2_approx(tr.x), ex2_approx(tr.y)));
                    const float2 sg = make_float2(rcp_approx(dr.x), rcp_approx(dr.y));
                    const float2 tn = __ffma2_rn(__fadd2_rn(make_float2(nh[j], nh[j + 1]), b_h), sg, __fadd2_rn(make_float2(nx[j], nx[j + 1]), b_x));
                    const float2 dn = __fadd2_rn(make_float2(ex2_approx(tn.x), ex2_approx(tn.y)), one2);
                    const float2 nn = __ffma2_rn(minus2, make_float2(rcp_approx(dn.x), rcp_approx(dn.y)), one2);
                    nx[j] = nn.x; nx[j + 1] = nn.y;
                }
            }
            else{
#pragma unroll
                for(int j = 0; j < HD; j++) nx[j] = tanh_of_scaled((nx[j] + bias[32 + j]) + (nh[j] + bias[48 + j]) * sigmoid_of_scaled(r[j] + bias[j]));   // scaled image
            }
            if constexpr(!B200L2F_TMEM_PREFETCH){
                tc::tmem_ld16(tmem_base + lane_off + C_D2 + 16, z);
                if constexpr(!B200L2F_H_IN_SMEM){
                    tc::tmem_ld16(tmem_base + lane_off + C_H_HI, hh);
                    tc::tmem_ld16(tmem_base + lane_off + C_H_LO, hl);
                }
            }
            if constexpr(B200L2F_H_IN_SMEM){       // h itself (hi + lo == h exactly, so the bits are those of the TMEM read-back)
                load_h(hh);
#pragma unroll
                for(int j = 0; j < HD; j++) hl[j] = 0.0f;
            }
            tc::tmem_ld_wait();
            if constexpr(B200L2F_GATES_PACKED){
                const float2 one2 = make_float2(1.0f, 1.0f);
#pragma unroll
                for(int j = 0; j < HD; j += 2){
                    const float2 b_z = *reinterpret_cast<const float2*>(bias + 16 + j);
                    const float2 tz = __fadd2_rn(make_float2(z[j], z[j + 1]), b_z);
                    const float2 dz = __fadd2_rn(one2, make_float2(ex2_approx(tz.x), ex2_approx(tz.y)));
                    const float2 zz = make_float2(rcp_approx(dz.x), rcp_approx(dz.y));
                    const float2 n2 = make_float2(nx[j], nx[j + 1]);
                    const float2 h2 = __fadd2_rn(make_float2(hh[j], hh[j + 1]), make_float2(hl[j], hl[j + 1]));
                    const float2 o = __ffma2_rn(zz, __fadd2_rn(h2, make_float2(-n2.x, -n2.y)), n2);   // (1 - z) n + z h
                    hn[j] = o.x; hn[j + 1] = o.y;
                }
            }
            else{
#pragma unroll
                for(int j = 0; j < HD; j++){ const float zz = sigmoid_of_scaled(z[j] + bias[16 + j]); hn[j] = fmaf(zz, (hh[j] + hl[j]) - nx[j], nx[j]); }   // (1 - z) n + z h
            }
        }
        // ---- dense 2 (16 -> 4) on the CUDA cores
        float act[4];
        {
            const float* w2 = sm_b + TcImage::W2T;
            const float4 b = *reinterpret_cast<const float4*>(w2 + 64);
            if constexpr(B200L2F_DENSE2_PACKED){   // same FMAs in the same order, two outputs per instruction
                float2 a01 = make_float2(b.x, b.y), a23 = make_float2(b.z, b.w);
#pragma unroll
                for(int k = 0; k < HD; k++){
                    const float4 w = *reinterpret_cast<const float4*>(w2 + 4 * k);
                    const float2 hk = make_float2(hn[k], hn[k]);
                    a01 = __ffma2_rn(make_float2(w.x, w.y), hk, a01); a23 = __ffma2_rn(make_float2(w.z, w.w), hk, a23);
                }
                act[0] = a01.x; act[1] = a01.y; act[2] = a23.x; act[3] = a23.y;
            }
            else{
            act[0] = b.x; act[1] = b.y; act[2] = b.z; act[3] = b.w;
#pragma unroll
            for(int k = 0; k < HD; k++){
                const float4 w = *reinterpret_cast<const float4*>(w2 + 4 * k);
                act[0] += w.x * hn[k]; act[1] += w.y * hn[k]; act[2] += w.z * hn[k]; act[3] += w.w * hn[k];
            }
            }
        }
        {   // gru/operations_generic.h:400-410: this step's output is kept, the stored state resets when the counter wraps
            const int new_step = gs + 1;
            const bool wrap = !no_auto_reset && new_step >= a.seq_len;
            if(wrap){ put8(C_H_HI, C_H_LO, sm_b + TcImage::H0); put8(C_H_HI + 8, C_H_LO + 8, sm_b + TcImage::H0 + 8); keep_h(sm_b + TcImage::H0); }
            else{ put8p(C_H_HI, C_H_LO, hn); put8p(C_H_HI + 8, C_H_LO + 8, hn + 8); keep_h(hn); }
            gs = wrap ? 0 : new_step;
        }
        if(RECORD && a.out_actions && active) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env) * 4) = make_float4(act[0], act[1], act[2], act[3]);
        RewardInputs ri;
        reward_inputs(ri, st);
        if constexpr(PIPELINE_G1){
            if(Spec::H == 1 || active) env_step_compiled<Spec, false, NOISE, true, AXIAL, false>(st, p, d, act, rng, hist_ptr, n);
            if(t + 1 < t_end) start_g1(t + 1);                 // the next step's dense 1 is on the tensor core from here on
            if(Spec::H == 1 || active) langevin_update_compiled<Spec, true>(st, p, rng, d.dt);
        }
        else{
            if(Spec::H == 1 || active) env_step_compiled<Spec, false, NOISE, true, AXIAL>(st, p, d, act, rng, hist_ptr, n, HOIST ? lang_normals : nullptr);
        }
        const bool term = env_terminated(p, st.x);
        const float rw = env_reward<true>(p, ri, act, st.x, term, d.dt);
        if(RECORD && a.out_rewards && active) a.out_rewards[(size_t)t * n + env] = rw;
        if(RECORD && a.out_term && active) a.out_term[(size_t)t * n + env] = term ? 1 : 0;
        if(!done){ ret += rw; eplen += 1; done = term; }
    }
    tc::tmem_st_wait();
    float hh[HD], hl[HD];   // tcgen05.ld is warp-collective (.sync.aligned): every lane executes it, only active lanes store
    if constexpr(B200L2F_H_IN_SMEM){
        load_h(hh);
#pragma unroll
        for(int j = 0; j < HD; j++) hl[j] = 0.0f;
    }
    else{
        tc::tmem_ld16(tmem_base + lane_off + C_H_HI, hh);
        tc::tmem_ld16(tmem_base + lane_off + C_H_LO, hl);
        tc::tmem_ld_wait();
    }
    const bool last_chunk = chunk == n_chunks - 1;
    if(active){
        if(RECORD && last_chunk && a.out_states && (a.T % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(a.T / a.state_stride) * n + env) * Spec::STATE_DIM);
        store_state(st, a.state + env, n);
        a.rng[env] = rng;
#pragma unroll
        for(int j = 0; j < HD; j++) a.hidden[(size_t)j * n + env] = hh[j] + hl[j];
        a.gru_step[env] = gs;
        if(last_chunk){
            if(a.out_returns) a.out_returns[env] = ret;
            if(a.out_eplen) a.out_eplen[env] = eplen;
            if(a.out_done) a.out_done[env] = done ? 1 : 0;
        }
        else{ a.acc_ret[env] = ret; a.acc_len[env] = (eplen << 1) | (done ? 1 : 0); }
    }
    if(!last_chunk){   // publish: every thread's stores, then the tile's progress counter
        __threadfence();
        __syncthreads();
        if(tid == 0) atomicExch(a.sched + 1 + tile, chunk + 1);
    }
    }   // work loop
    tc::tc_fence_before();
    __syncthreads();
    if(warp == 0) tc::tmem_dealloc<128>(tmem_base);
}

}  // namespace b200l2f
