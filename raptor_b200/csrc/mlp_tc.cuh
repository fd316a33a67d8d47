// raptor_b200/csrc/mlp_tc.cuh -- MLP actors (SAC teacher / PPO, hidden 64) on the sm_100a tensor cores, for the H = 1 specs (RAPTOR, TEACHER):
//   k_rollout_mlp_ts   closed-loop rollout with a deterministic MLP actor          (BASELINE config 3; CUDA-core twin: k_rollout_mlp)
//   k_collect_ts       PPO collection with on-device auto-reset and write-back     (BASELINE config 4; CUDA-core twin: k_collect)
// Reference semantics are those of mlp.cuh (INC/nn_models/mlp/network.h:15-51, standardize, squash, on_policy_runner) -- only the three dense
// layers move: each is one [128 envs] x [K] x [N] GEMM per CTA and step, tcgen05.mma kind::tf32 with the 3xTF32 split
// (A_hi B_hi + A_hi B_lo + A_lo B_hi, fp32 accumulate), the A operand (activations, lane = environment) and the accumulators in tensor
// memory, the weight image (hi/lo planes, canonical K-major core-matrix order) in shared memory, loaded once per CTA by a TMA bulk copy.
//
// TMEM plan (256 columns per CTA, 2 CTAs/SM):      layer 1: A1 hi [0,32)   lo [32,64)    -> D1 [64,128)
//                                                  layer 2: A2 hi [128,192) lo [192,256) -> D2 [0,64)      (A1 is dead)
//                                                  layer 3: A3 hi [64,128)  lo [128,192) -> D3 [192,208)   (D1, A2 are dead)
// Wide first layer (K1 > 32: the DEFAULT spec's 82-wide observation, K1 = 88, eleven K = 8 operand blocks per plane):
//                                                  layer 1: A1 hi [0,88)   lo [88,176)   -> D1 [176,240)
//                                                  layer 2: A2 hi [0,64)   lo [64,128)   -> D2 [128,192)   (A1 dead; D1 [176,192) is read before G2 is issued)
//                                                  layer 3: A3 hi [0,64)   lo [64,128)   -> D3 [240,256)   (A2, D1 dead)
// Layer 1 carries its bias as an extra K column (A = 1.0); the biases of layers 2 and 3 (K = 64 is full) are added in the epilogues.
#pragma once
#include "mlp.cuh"
#include "rollout_tc.cuh"

#ifndef B200L2F_L3_CUDA
#define B200L2F_L3_CUDA 1               // actors / critics with <= 4 outputs: last MLP layer on the CUDA cores (one tensor-core round trip less per step)
#endif
#ifndef B200L2F_L3_CUDA8
#define B200L2F_L3_CUDA8 1              // the same for eight outputs (SAC actors): +2.5 % on config 3, profiles/r02_exp10_last_layer_cuda_cores.log
#endif
#ifndef B200L2F_FAST_RESET
#define B200L2F_FAST_RESET 1            // in-kernel resets of the default-math collection / runner kernels on the MUFU pipe (samplers.cuh: FAST twins)
#endif
#ifndef B200L2F_COLLECT_ROLLED_RK4
#define B200L2F_COLLECT_ROLLED_RK4 1   // the collection kernel's integrator as one rolled stage loop (code size, see k_collect_ts)
#endif

namespace b200l2f {

template <int OUT> constexpr bool mlp_l3_cuda_default(){ return OUT <= 4 ? B200L2F_L3_CUDA != 0 : B200L2F_L3_CUDA8 != 0; }

template <int IN, int OUT>
struct MlpTcImage {
    static constexpr int HD = MLP_HD;
    static constexpr int K1 = (IN + 1 + 7) / 8 * 8;   // observation + bias column, padded to the instruction K
    static constexpr int N3 = 16;                     // smallest N of an M = 128 instruction
    static_assert(K1 <= 88 && OUT <= N3, "TMEM plan");
    static constexpr bool WIDE = K1 > 32;             // which TMEM column plan mlp_forward_ts uses (see the file header)
    // L3_CUDA: the last layer runs on the CUDA cores (mlp_forward_ts_from) from an fp32 k-major copy of W3 -- the image then carries that copy INSTEAD of the hi / lo
    // operand planes of layer 3 (7 / 6 KB less shared memory per CTA, which is what keeps the runner kernels at two CTAs per SM)
    static constexpr bool L3_CUDA = mlp_l3_cuda_default<OUT>();
    static constexpr int B1_HI = 0, B1_LO = B1_HI + K1 * HD;
    static constexpr int B2_HI = B1_LO + K1 * HD, B2_LO = B2_HI + HD * HD;
    static constexpr int B3_HI = B2_LO + HD * HD, B3_LO = B3_HI + (L3_CUDA ? 0 : HD * N3);
    static constexpr int MEAN = B3_LO + (L3_CUDA ? 0 : HD * N3), PREC = MEAN + K1;
    static constexpr int BIAS2 = PREC + K1, BIAS3 = BIAS2 + HD, LOG_STD = BIAS3 + N3;
    static constexpr int W3T = LOG_STD + 4;           // L3_CUDA: the last layer in fp32, k-major [64][W3N], times 0.5 (the hidden activations arrive doubled)
    static constexpr int W3N = L3_CUDA ? (OUT <= 4 ? 4 : 8) : 0;
    static constexpr int SIZE = W3T + HD * W3N;
    static constexpr int BYTES = SIZE * 4;
    static_assert(BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
};
// blob (include/b200_l2f.h MLP order; standardize / log_std blocks optional) -> image
template <int IN, int OUT>
inline void build_mlp_tc_image_host(float* img, const float* blob, bool has_std, bool has_log_std){
    using I = MlpTcImage<IN, OUT>;
    constexpr int HD = MLP_HD;
    for(int i = 0; i < I::SIZE; i++) img[i] = 0.0f;
    const float* b = blob;
    for(int i = 0; i < IN; i++){ img[I::MEAN + i] = has_std ? b[i] : 0.0f; img[I::PREC + i] = has_std ? b[IN + i] : 1.0f; }
    if(has_std) b += 2 * IN;
    const float* W1 = b; const float* b1 = W1 + HD * IN; const float* W2 = b1 + HD; const float* b2 = W2 + HD * HD;
    const float* W3 = b2 + HD; const float* b3 = W3 + OUT * HD; const float* ls = b3 + OUT;
    auto put = [&](int hi_base, int lo_base, int N, int n, int k, float v){
        float hi, lo; tc_split_host(v, hi, lo);
        const int idx = (k / 4) * N * 4 + n * 4 + (k % 4);
        img[hi_base + idx] = hi; img[lo_base + idx] = lo;
    };
    for(int n = 0; n < HD; n++){
        for(int k = 0; k < IN; k++) put(I::B1_HI, I::B1_LO, HD, n, k, W1[n * IN + k]);
        put(I::B1_HI, I::B1_LO, HD, n, IN, b1[n]);
        for(int k = 0; k < HD; k++) put(I::B2_HI, I::B2_LO, HD, n, k, 0.5f * W2[n * HD + k]);   // the hidden activations arrive doubled (packed ReLU, see relu2x): exact
        img[I::BIAS2 + n] = b2[n];
    }
    for(int n = 0; n < OUT; n++){
        if constexpr(!I::L3_CUDA){ for(int k = 0; k < HD; k++) put(I::B3_HI, I::B3_LO, I::N3, n, k, 0.5f * W3[n * HD + k]); }
        img[I::BIAS3 + n] = b3[n];
    }
    for(int i = 0; i < 4; i++) img[I::LOG_STD + i] = has_log_std ? ls[i] : 0.0f;
    for(int k = 0; k < HD; k++) for(int n = 0; n < I::W3N; n++) img[I::W3T + I::W3N * k + n] = n < OUT ? 0.5f * W3[n * HD + k] : 0.0f;
}

// per-thread view of the CTA's tensor-core state
struct TsCtx {
    const float* sm_b;        // weight image in shared memory
    uint32_t b_s;             // its shared-memory address
    uint32_t tmem_base;       // allocation base (lane 0)
    uint32_t tmem_lane;       // base + this warp's lane offset: what tcgen05.ld / st address
    uint64_t* bar_mma;
    uint32_t phase;
    int tid;
    const volatile int* probe; int probe_value;   // see ts_run<.., PROBE>
    uint64_t desc_n64, desc_n16;   // B descriptors of the image base for N = 64 / N = 16 tiles: every other descriptor is a compile-time offset away
};
__device__ __forceinline__ void ts_put8(uint32_t t_hi, uint32_t t_lo, const float* v){
    float hi[8], lo[8];
#pragma unroll
    for(int i = 0; i < 8; i += 2){                        // hi: one LOP3 each; lo = v - hi (exact) two at a time on the packed pipe
        hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u); hi[i + 1] = __uint_as_float(__float_as_uint(v[i + 1]) & 0xFFFFE000u);
        const float2 l = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(-hi[i], -hi[i + 1]));
        lo[i] = l.x; lo[i + 1] = l.y;
    }
    tc::tmem_st8(t_hi, hi);
    tc::tmem_st8(t_lo, lo);
}
// 2 ReLU(v) = v + |v| on the packed pipe: one FADD2 (|.| is an operand modifier) for two activations instead of two FMNMX.  The factor 2 is exact (an exponent
// increment, carried through the hi / lo split) and is taken back by the 0.5 folded into the next layer's weights in the image: bit-identical to max(v, 0).
__device__ __forceinline__ void relu2x_pairs(float* __restrict__ v, int n){
#pragma unroll
    for(int j = 0; j < n; j += 2){
        const float2 a = make_float2(v[j], v[j + 1]);
        const float2 r = __fadd2_rn(a, make_float2(fabsf(a.x), fabsf(a.y)));
        v[j] = r.x; v[j + 1] = r.y;
    }
}
// ksteps instructions of K = 8: A columns [a_hi + 8 s, +8) / [a_lo + 8 s, +8) against the B chunk pair s (issued by one thread)
__device__ __forceinline__ void ts_issue_gemm(const TsCtx& c, uint32_t dcol, uint32_t a_hi, uint32_t a_lo, int ksteps, int b_hi_off, int b_lo_off, uint32_t N, uint32_t idesc){
    uint32_t acc = 0;
    const uint64_t base = N == 16 ? c.desc_n16 : c.desc_n64;
#pragma unroll
    for(int s = 0; s < ksteps; s++){
        const uint64_t bhi = tc::smem_desc_advance(base, b_hi_off * 4 + s * 2 * N * 16);
        const uint64_t blo = tc::smem_desc_advance(base, b_lo_off * 4 + s * 2 * N * 16);
        tc::mma_tf32_ts(c.tmem_base + dcol, c.tmem_base + a_hi + 8 * s, bhi, idesc, acc); acc = 1;
        tc::mma_tf32_ts(c.tmem_base + dcol, c.tmem_base + a_hi + 8 * s, blo, idesc, 1);
        tc::mma_tf32_ts(c.tmem_base + dcol, c.tmem_base + a_lo + 8 * s, bhi, idesc, 1);
    }
}
// the A operand of the next GEMM is complete in TMEM: hand it to the tensor core, wait for the accumulator.  Called by ALL threads of the CTA.
// NAMED_BAR: the CTA holds more warps than the four that own the tile (k_collect_lag's reset warp): those four meet on barrier 1
// PROBE: read *c.probe right behind the barrier (a CTA-uniform value: every warp reads it between the same two barriers)
template <bool NAMED_BAR = false, bool PROBE = false, class F>
__device__ __forceinline__ void ts_run(TsCtx& c, F&& issue){
    tc::tmem_st_wait();
    tc::tc_fence_before();
    if constexpr(NAMED_BAR) asm volatile("bar.sync 1, 128;" ::: "memory");
    else __syncthreads();
    if constexpr(PROBE) c.probe_value = *c.probe;
    if((c.tid >> 5) == 0 && tc::elect_one()){   // one lane of warp 0 (elect.sync): straight-line predicated MMA issue
        tc::tc_fence_after();
        issue();
        tc::mma_commit(c.bar_mma);
    }
    tc::mbar_wait(c.bar_mma, c.phase); c.phase ^= 1;
    tc::tc_fence_after();
}

// obs_at(k): RAW observation column k of this thread's environment (k is a compile-time constant after unrolling: registers or shared memory)
// -> out[OUT] (pre-head outputs).  Called by all 128 threads (inactive lanes compute garbage rows).
// L3_CUDA (OUT <= 4): the last layer (64 -> OUT) on the CUDA cores, straight from the registers that hold the second hidden layer: 2 x 64 FFMA2 + 64 broadcast LDS.128
// instead of splitting / storing the 64 activations to TMEM and a third barrier + MMA round trip (24 N = 16 instructions for at most 4 useful columns)
template <int IN, int OUT, bool NAMED_BAR = false, class OBS>
__device__ __forceinline__ void mlp_forward_ts_from(TsCtx& c, OBS&& obs_at, float* __restrict__ out){
    using I = MlpTcImage<IN, OUT>;
    constexpr bool L3_CUDA = I::L3_CUDA;
    constexpr int HD = MLP_HD, K1 = I::K1;
    constexpr bool WIDE = I::WIDE;
    constexpr uint32_t A1_HI = 0, A1_LO = WIDE ? 88 : 32, D1 = WIDE ? 176 : 64, A2_HI = WIDE ? 0 : 128, A2_LO = WIDE ? 64 : 192, D2 = WIDE ? 128 : 0,
                       A3_HI = WIDE ? 0 : 64, A3_LO = WIDE ? 64 : 128, D3 = WIDE ? 240 : 192;
    constexpr uint32_t IDESC64 = tc::make_idesc_tf32(128, 64), IDESC16 = tc::make_idesc_tf32(128, 16);
#pragma unroll
    for(int g = 0; g < K1 / 8; g++){       // one K = 8 operand block at a time: at most 8 observation values live in registers
        float x[8];
#pragma unroll
        for(int j = 0; j < 8; j++){
            const int k = 8 * g + j;
            if(k < IN){                    // standardize: (x - mean) [* precision unless it is 0]
                float v = obs_at(k) - c.sm_b[I::MEAN + k];
                const float pr = c.sm_b[I::PREC + k];
                if(pr != 0.0f) v *= pr;
                x[j] = v;
            }
            else x[j] = k == IN ? 1.0f : 0.0f;   // bias column, padding
        }
        ts_put8(c.tmem_lane + A1_HI + 8 * g, c.tmem_lane + A1_LO + 8 * g, x);
    }
    ts_run<NAMED_BAR, NAMED_BAR>(c, [&](){ ts_issue_gemm(c, D1, A1_HI, A1_LO, K1 / 8, I::B1_HI, I::B1_LO, HD, IDESC64); });
#pragma unroll
    for(int g = 0; g < HD / 32; g++){
        float v[32];
        tc::tmem_ld16(c.tmem_lane + D1 + 32 * g, v);
        tc::tmem_ld16(c.tmem_lane + D1 + 32 * g + 16, v + 16);
        tc::tmem_ld_wait();
        relu2x_pairs(v, 32);
#pragma unroll
        for(int q = 0; q < 4; q++) ts_put8(c.tmem_lane + A2_HI + 32 * g + 8 * q, c.tmem_lane + A2_LO + 32 * g + 8 * q, v + 8 * q);
    }
    ts_run<NAMED_BAR>(c, [&](){ ts_issue_gemm(c, D2, A2_HI, A2_LO, HD / 8, I::B2_HI, I::B2_LO, HD, IDESC64); });
    float2 acc01 = make_float2(0.0f, 0.0f), acc23 = acc01, acc45 = acc01, acc67 = acc01;   // L3_CUDA: outputs 0 | 1, 2 | 3, 4 | 5, 6 | 7
#pragma unroll
    for(int g = 0; g < HD / 32; g++){
        float v[32];
        tc::tmem_ld16(c.tmem_lane + D2 + 32 * g, v);
        tc::tmem_ld16(c.tmem_lane + D2 + 32 * g + 16, v + 16);
        tc::tmem_ld_wait();
#pragma unroll
        for(int j4 = 0; j4 < 8; j4++){
            const float4 b = *reinterpret_cast<const float4*>(c.sm_b + I::BIAS2 + 32 * g + 4 * j4);
            const float2 s01 = __fadd2_rn(make_float2(v[4 * j4], v[4 * j4 + 1]), make_float2(b.x, b.y)), s23 = __fadd2_rn(make_float2(v[4 * j4 + 2], v[4 * j4 + 3]), make_float2(b.z, b.w));
            v[4 * j4] = s01.x; v[4 * j4 + 1] = s01.y; v[4 * j4 + 2] = s23.x; v[4 * j4 + 3] = s23.y;
        }
        relu2x_pairs(v, 32);
        if constexpr(L3_CUDA){
            if(g == 0){
                acc01 = make_float2(c.sm_b[I::BIAS3], c.sm_b[I::BIAS3 + 1]); acc23 = make_float2(c.sm_b[I::BIAS3 + 2], c.sm_b[I::BIAS3 + 3]);
                if constexpr(OUT > 4){ acc45 = make_float2(c.sm_b[I::BIAS3 + 4], c.sm_b[I::BIAS3 + 5]); acc67 = make_float2(c.sm_b[I::BIAS3 + 6], c.sm_b[I::BIAS3 + 7]); }
            }
#pragma unroll
            for(int k = 0; k < 32; k++){
                const float* wr = c.sm_b + I::W3T + I::W3N * (32 * g + k);                                // same address in every lane: broadcast
                const float4 w = *reinterpret_cast<const float4*>(wr);
                const float2 x = make_float2(v[k], v[k]);
                acc01 = __ffma2_rn(x, make_float2(w.x, w.y), acc01);
                if constexpr(OUT > 2) acc23 = __ffma2_rn(x, make_float2(w.z, w.w), acc23);
                if constexpr(OUT > 4){
                    const float4 u = *reinterpret_cast<const float4*>(wr + 4);
                    acc45 = __ffma2_rn(x, make_float2(u.x, u.y), acc45);
                    acc67 = __ffma2_rn(x, make_float2(u.z, u.w), acc67);
                }
            }
        }
        else{
#pragma unroll
            for(int q = 0; q < 4; q++) ts_put8(c.tmem_lane + A3_HI + 32 * g + 8 * q, c.tmem_lane + A3_LO + 32 * g + 8 * q, v + 8 * q);
        }
    }
    if constexpr(L3_CUDA){
        const float o8[8] = {acc01.x, acc01.y, acc23.x, acc23.y, acc45.x, acc45.y, acc67.x, acc67.y};
#pragma unroll
        for(int j = 0; j < OUT; j++) out[j] = o8[j];
    }
    else{
        ts_run<NAMED_BAR>(c, [&](){ ts_issue_gemm(c, D3, A3_HI, A3_LO, HD / 8, I::B3_HI, I::B3_LO, I::N3, IDESC16); });
        float v[16];
        tc::tmem_ld16(c.tmem_lane + D3, v);
        tc::tmem_ld_wait();
#pragma unroll
        for(int j = 0; j < OUT; j++) out[j] = v[j] + c.sm_b[I::BIAS3 + j];
    }
}
template <int IN, int OUT, bool NAMED_BAR = false>
__device__ __forceinline__ void mlp_forward_ts(TsCtx& c, const float* __restrict__ obs, float* __restrict__ out){   // observation in registers
    mlp_forward_ts_from<IN, OUT, NAMED_BAR>(c, [&](int k){ return obs[k]; }, out);
}

// full observation of an H = 1 spec in registers (same values and RNG order as observe_to_scratch)
template <class Spec, bool NOISE, bool FAST = false, class P>
__device__ __forceinline__ void observe_regs(const EnvState<Spec>& st, const P& p, uint64_t& rng, float* __restrict__ o){
    static_assert(Spec::H == 1, "register observation: H = 1 specs");
    observe18<Spec, NOISE, FAST>(st, p, rng, o);
#pragma unroll
    for(int i = 0; i < 4; i++) o[18 + i] = st.hist[i];
    if constexpr(Spec::OBS_LAYOUT == OBS_TEACHER){
#pragma unroll
        for(int i = 0; i < 4; i++) o[22 + i] = (st.x[X_RPM + i] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2.0f - 1.0f;
    }
}

template <int IN, int OUT>
struct MlpTsSmem {
    static constexpr int B = 0;
    static constexpr int DYN = (MlpTcImage<IN, OUT>::BYTES + 127) / 128 * 128;
    static constexpr int BAR = DYN + C_DIM * BLOCK * 4;
    static constexpr int SLAB = BAR + 32;                                 // collect only: per-warp [32][W] write-back windows
    static constexpr int TOTAL_ROLLOUT = SLAB;
    static constexpr int TOTAL_COLLECT = SLAB + 4 * 32 * (IN + 15) * 4;   // the warp's 32 dataset rows, row stride D = IN + 15 (odd for every spec: conflict-free)
    static constexpr int HIST_RING = 64 * BLOCK * 4;                      // H = 16 specs: the CTA's action-history rings [16 x 4][128] (k_collect_ts keeps them on chip)
};

// CTA prologue shared by both kernels: barriers, TMEM allocation, weight image by TMA.  Returns the context; every thread must call it.
template <int IN, int OUT>
__device__ __forceinline__ TsCtx mlp_ts_prologue_at(unsigned char* smraw, int b_offset, int bar_offset, const float* __restrict__ tc_image){
    float* sm_b = reinterpret_cast<float*>(smraw + b_offset);
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smraw + bar_offset);
    uint64_t* bar_mma = bar_tma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
    const int tid = threadIdx.x, warp = tc::uniform_warp_index();
    if(tid == 0){
        tc::mbar_init(bar_tma, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_fence_init();
    }
    if(warp == 0) tc::tmem_alloc<256>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    TsCtx c;
    c.sm_b = sm_b; c.b_s = tc::smem_u32(sm_b); c.tmem_base = *tmem_slot; c.tmem_lane = c.tmem_base + ((uint32_t)(warp * 32) << 16);
    c.bar_mma = bar_mma; c.phase = 0; c.tid = tid;
    c.desc_n64 = tc::make_smem_desc(c.b_s, 64 * 16, 128); c.desc_n16 = tc::make_smem_desc(c.b_s, 16 * 16, 128);
    if(tc_image == nullptr) return c;          // the caller loads (and reloads) the image itself (dagger.cuh: one image per teacher)
    if(tid == 0){
        tc::mbar_expect_tx(bar_tma, MlpTcImage<IN, OUT>::BYTES);
        tc::tma_load_1d(sm_b, tc_image, MlpTcImage<IN, OUT>::BYTES, bar_tma);
    }
    tc::mbar_wait(bar_tma, 0);
    __syncthreads();
    return c;
}
template <int IN, int OUT>
__device__ __forceinline__ TsCtx mlp_ts_prologue(unsigned char* smraw, const float* __restrict__ tc_image){
    return mlp_ts_prologue_at<IN, OUT>(smraw, MlpTsSmem<IN, OUT>::B, MlpTsSmem<IN, OUT>::BAR, tc_image);
}
__device__ __forceinline__ void mlp_ts_epilogue(const TsCtx& c){
    tc::tc_fence_before();
    __syncthreads();
    if((c.tid >> 5) == 0) tc::tmem_dealloc<256>(c.tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
// closed-loop rollout, deterministic MLP actor (head identity or squash-eval), persistent (tile, time-chunk) work queue as in
// k_rollout_raptor_ts (the actor is stateless, so a chunk hands over the environment state and the episode accumulators only)
// ---------------------------------------------------------------------------------------------------------------
template <class Spec, int OUT, bool UNIFORM, bool AXIAL, bool NOISE = false>
__global__ void __launch_bounds__(BLOCK, 2) k_rollout_mlp_ts(const __grid_constant__ RolloutArgs a, const float* __restrict__ tc_image){
    constexpr int IN = Spec::OBS_DIM;
    using SM = MlpTsSmem<IN, OUT>;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* sm_dyn = reinterpret_cast<float*>(smraw + SM::DYN);
    TsCtx c = mlp_ts_prologue<IN, OUT>(smraw, tc_image);
    const int tid = threadIdx.x;
    const size_t n = (size_t)a.n;
    __shared__ int s_item;
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
        __threadfence();
    }
    const int t_begin = chunk * a.chunk_steps;
    const int t_end = min(a.T, t_begin + a.chunk_steps);
    const int e = tile * BLOCK + tid;
    const bool active = e < a.n;
    const size_t env = active ? (size_t)e : 0;
    const ParamsCompiledT<UNIFORM> p = stage_dynamics_compiled<UNIFORM>(sm_dyn, a.params, n, env, a.row0);
    EnvState<Spec> st;
    load_state_cg(st, a.state + env, n);
    DynInvariants d;
    {
        ParamsGlobal pg{a.params + env, n};
        dyn_invariants(d, pg, st);
    }
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = __ldcg(a.rng + env);
    float ret = 0.0f; int eplen = 0; bool done = false;
    if(chunk > 0){ ret = __ldcg(a.acc_ret + env); const int v = __ldcg(a.acc_len + env); eplen = v >> 1; done = (v & 1) != 0; }

    for(int t = t_begin; t < t_end; t++){
        if(a.out_states && active && (t % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(t / a.state_stride) * n + env) * Spec::STATE_DIM);
        float obs[IN];
        observe_regs<Spec, NOISE, true>(st, p, rng, obs);
        if(a.out_obs && active){
            float* row = a.out_obs + ((size_t)t * n + env) * IN;
#pragma unroll
            for(int i = 0; i < IN; i++) row[i] = obs[i];
        }
        float o[OUT], act[4];
        mlp_forward_ts<IN, OUT>(c, obs, o);
#pragma unroll
        for(int i = 0; i < 4; i++) act[i] = OUT == 8 ? tanhf(o[i]) : o[i];
        if(a.out_actions && active) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env) * 4) = make_float4(act[0], act[1], act[2], act[3]);
        RewardInputs ri;
        reward_inputs(ri, st);
        env_step_compiled<Spec, false, NOISE, true, AXIAL>(st, p, d, act, rng, hist_ptr, n);
        const bool term = env_terminated(p, st.x);
        const float rw = env_reward<true>(p, ri, act, st.x, term, d.dt);
        if(a.out_rewards && active) a.out_rewards[(size_t)t * n + env] = rw;
        if(a.out_term && active) a.out_term[(size_t)t * n + env] = term ? 1 : 0;
        if(!done){ ret += rw; eplen += 1; done = term; }
    }
    const bool last_chunk = chunk == n_chunks - 1;
    if(active){
        if(last_chunk && a.out_states && (a.T % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(a.T / a.state_stride) * n + env) * Spec::STATE_DIM);
        store_state(st, a.state + env, n);
        a.rng[env] = rng;
        if(last_chunk){
            if(a.out_returns) a.out_returns[env] = ret;
            if(a.out_eplen) a.out_eplen[env] = eplen;
            if(a.out_done) a.out_done[env] = done ? 1 : 0;
        }
        else{ a.acc_ret[env] = ret; a.acc_len[env] = (eplen << 1) | (done ? 1 : 0); }
    }
    if(!last_chunk){
        __threadfence();
        __syncthreads();
        if(tid == 0) atomicExch(a.sched + 1 + tile, chunk + 1);
    }
    }   // work loop
    mlp_ts_epilogue(c);
}

// ---------------------------------------------------------------------------------------------------------------
// PPO collection (rl_tools::collect), same dataset contract as k_collect.  One tile per CTA visit (persistent tile loop); an environment's
// reset (parameter / state re-sampling) is divergent CUDA-core work inside the step loop, the three GEMMs are CTA-collective and therefore
// sit outside every lane-dependent branch.
// ---------------------------------------------------------------------------------------------------------------
// FOLLOW: every parameter column was filled from a.row before (initial_parameters / sample_initial_parameters / earlier resets) -- the
// non-randomised entries equal the row, so the MDP constants are read from the constant bank and a reset writes only the randomised entries.
// The RK4 stages are one rolled loop here: next to the reset samplers the straight-line integrator would not fit the instruction cache.
template <class Spec, bool DR, bool FOLLOW, bool AXIAL>
__global__ void __launch_bounds__(BLOCK, 2) k_collect_ts(const __grid_constant__ CollectArgs a, const float* __restrict__ tc_image, int* __restrict__ sched){
    constexpr int IN = Spec::OBS_DIM, OUT = 4;
    constexpr int D = IN + 15, W = IN + 12, WS = D;        // W columns produced per step; the staging window has the dataset's own row stride
    using SM = MlpTsSmem<IN, OUT>;
    using I = MlpTcImage<IN, OUT>;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* sm_dyn = reinterpret_cast<float*>(smraw + SM::DYN);
    TsCtx c = mlp_ts_prologue<IN, OUT>(smraw, tc_image);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* slab = reinterpret_cast<float*>(smraw + SM::SLAB) + (size_t)warp * 32 * WS;   // private to this warp: the image of its 32 consecutive dataset rows
    float* myrow = slab + lane * WS;                       // odd stride: conflict-free
#pragma unroll
    for(int i = W; i < D; i++) myrow[i] = 0.0f;            // the learner's columns (value | advantage | target_value) leave as zeros, see b200l2f_collect
    const size_t n = (size_t)a.n;
    __shared__ int s_item;
    const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
    bool store_pending = false;                            // this warp has a bulk store in flight that still reads the window
    // Time-chunked scheduler (a.n_chunks > 1; as in k_rollout_mlp_ts): the collection of a tile is cut into chunks handed out chunk-major by one atomic counter, so
    // a tile count that is not a multiple of the resident CTAs (DEFAULT spec: one CTA per SM, 512 tiles on 148 SMs = 4 rounds for 3.46 rounds of work) no longer
    // quantises the launch.  A chunk hands the environment over through HBM exactly like the end of a launch does (state, RNG, runner bookkeeping, parameters); the
    // next chunk may run on another SM, so every global load of such a build must bypass L1 (the launcher enables chunks only for units compiled with -dlcm=cg).
    const int n_chunks = a.n_chunks > 1 ? a.n_chunks : 1;
    const int total_items = n_tiles * n_chunks;
    for(;;){
    if(tid == 0) s_item = atomicAdd(sched, 1);
    __syncthreads();
    const int item = s_item;
    __syncthreads();
    if(item >= total_items) break;
    const int tile = item % n_tiles, chunk = item / n_tiles;
    if(chunk > 0){
        if(tid == 0){
            const int* prog = sched + 1 + tile;
            int v;
            do{ asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(prog) : "memory"); if(v < chunk) __nanosleep(64); } while(v < chunk);
        }
        __syncthreads();
        __threadfence();
    }
    const int t_begin = n_chunks > 1 ? chunk * a.chunk_steps : 0;
    const int t_end = (chunk == n_chunks - 1) ? a.T + 1 : min(a.T + 1, t_begin + a.chunk_steps);   // the last chunk carries the final-observation row (t == T)
    const int e = tile * BLOCK + tid;
    const bool active = e < a.n;
    const size_t env = active ? (size_t)e : 0;
    ParamsCompiledT<FOLLOW, false, FOLLOW> p = stage_dynamics_compiled<FOLLOW, false, FOLLOW>(sm_dyn, a.params, n, env, a.row);
    EnvState<Spec> st;
    load_state_cg(st, a.state + env, n);                  // through L2: with time chunks another SM may have written it in this launch
    DynInvariants d;
    {
        ParamsRW pg{a.params + env, n};
        dyn_invariants(d, pg, st);
    }
    // H > 1 (DEFAULT spec): the 16-deep action-history ring of every environment lives in shared memory for the tile visit (element (h, a) of thread t at
    // ring[(4 h + a) * BLOCK + t]: the same (pointer, stride) form the environment code uses for the HBM rows, bank-conflict-free) -- the observation reads all 64 values
    // every step, which from HBM was one L2 round trip per step with nothing to hide it at one warp per scheduler
    float* const hist_hbm = a.state + (size_t)S_HIST * n + env;
    float* hist_ptr = hist_hbm;
    size_t hist_stride = n;
    if constexpr(Spec::H > 1){
        float* ring = reinterpret_cast<float*>(smraw + SM::TOTAL_COLLECT) + tid;
#pragma unroll 8
        for(int i = 0; i < 4 * Spec::H; i++) ring[i * BLOCK] = __ldcg(hist_hbm + (size_t)i * n);
        hist_ptr = ring; hist_stride = BLOCK;
    }
    uint64_t rng = __ldcg(a.rng + env);
    int ep_step = __ldcg(a.episode_step + env); float ep_ret = __ldcg(a.episode_return + env); bool truncated = __ldcg(a.truncated + env) != 0;
    uint64_t rng_at_reset = 0; bool reset_seen = false;
    const int warp_env0 = tile * BLOCK + warp * 32;
    const int rows_valid = min(32, a.n - warp_env0);
    // a full warp's 32 rows are one contiguous, 16-byte aligned run of 32 * D floats in the dataset: ONE bulk copy (TMA, shared -> global) per
    // warp and step instead of a store loop.  Ragged last warp / unaligned dataset: the element loop below.
    const bool bulk = a.bulk_rows != 0 && rows_valid == 32;

    for(int t = t_begin; t < t_end; t++){
        const bool last = t == a.T;                       // final observation only (operations_generic.h:122-129)
        if(!last && truncated && active){                 // prologue (operations_generic_per_env.h:17-25): re-sample parameters and state
            truncated = false; ep_step = 0; ep_ret = 0.0f;
            ParamsOverlay o;                              // sampled in registers: no dependent HBM round trips on the reset path
            o.init(a.row);
            if constexpr(DR && FOLLOW){ rng_at_reset = rng; reset_seen = true; }   // the column is written once, after the step loop (see there)
            if(!sample_parameters<DR, Spec::RNG_OOL, B200L2F_FAST_RESET != 0>(o, rng)) atomicExch(a.error_flag, 1);
            if constexpr(!FOLLOW) o.template flush<true>(ParamsRW{a.params + env, n});
            compile_dynamics_block<true, B200L2F_FAST_RESET != 0>(dyn_block_of_thread(sm_dyn), [&](int i){ return o[i]; });   // this thread's block only
            sample_state<Spec, ParamsOverlay, true, B200L2F_FAST_RESET != 0>(st, o, rng, hist_ptr, hist_stride);
            dyn_invariants<Spec, ParamsOverlay, B200L2F_FAST_RESET != 0>(d, o, st);
        }
        // the previous step's bulk store has read the window (in flight since the end of that step: the wait is free)
        if(store_pending){ if(lane == 0) tc::bulk_store_wait_read(); store_pending = false; }
        __syncwarp();
        // The observation goes straight into this lane's row of the window.  H == 1: it also stays in registers for the first layer's operand.
        // H > 1 (DEFAULT spec, 82 columns incl. the 16-deep action ring in HBM): the first layer's operand blocks are read back from the window,
        // eight columns at a time
        float obs[Spec::H == 1 ? IN : 1];
        if constexpr(Spec::H == 1){
            observe_regs<Spec, true, true>(st, p, rng, obs);
#pragma unroll
            for(int i = 0; i < IN; i++) myrow[i] = obs[i];
        }
        else observe_to_scratch<Spec, true>(st, p, rng, hist_ptr, hist_stride, myrow, 1);
        float vals[12];
#pragma unroll
        for(int i = 0; i < 12; i++) vals[i] = 0.0f;       // the final rows carry the observation only
        if(!last){                                        // uniform across the CTA
            float mean[OUT], act[4];
            if constexpr(Spec::H == 1) mlp_forward_ts<IN, OUT>(c, obs, mean);
            else mlp_forward_ts_from<IN, OUT>(c, [&](int k){ return myrow[k]; }, mean);
            float lp = 0.0f;
#pragma unroll
            for(int i = 0; i < 4; i++){                   // epilogue (operations_generic_per_env.h:43-58)
                const float ls = c.sm_b[I::LOG_STD + i];
                act[i] = rng_normal_t<Spec::RNG_OOL, true>(rng, mean[i], ex2_approx(ls * LOG2E));   // default math: MUFU Box-Muller / exponential (the integer stream stays bit-exact)
                lp += normal_log_prob(mean[i], ls, act[i]);
            }
            RewardInputs ri;
            reward_inputs(ri, st);
            env_step_compiled<Spec, B200L2F_COLLECT_ROLLED_RK4 != 0, true, true, AXIAL>(st, p, d, act, rng, hist_ptr, hist_stride);   // H > 1: the ring is this thread's own shared-memory column (shadow lanes too)
            const bool term = env_terminated(p, st.x);
            const float r = env_reward<true>(p, ri, act, st.x, term, d.dt);
            ep_ret += r; ep_step += 1;
            truncated = term || (a.step_limit > 0 && ep_step >= a.step_limit);
#pragma unroll
            for(int i = 0; i < 4; i++){ vals[i] = mean[i]; vals[4 + i] = act[i]; }
            vals[8] = lp; vals[9] = r; vals[10] = term ? 1.0f : 0.0f; vals[11] = truncated ? 1.0f : 0.0f;
        }
#pragma unroll
        for(int i = 0; i < 12; i++) myrow[IN + i] = vals[i];
        float* gbase = a.dataset + ((size_t)t * n + warp_env0) * D;
        if(bulk){
            tc::fence_async_smem();                       // this lane's row -> visible to the async proxy
            __syncwarp();
            if(lane == 0) tc::bulk_store(gbase, slab, 32 * D * 4);
            store_pending = true;
        }
        else{
            __syncwarp();
            // the window is the rows back to back (learner columns included: zeros), so is the dataset
#pragma unroll 2
            for(int it = 0; it < D; it++){
                const int idx = lane + 32 * it;
                const int r = idx / D;
                if(r < rows_valid) gbase[idx] = slab[idx];
            }
            __syncwarp();
        }
    }
    if(active){
        store_state(st, a.state + env, n);
        a.rng[env] = rng;
        a.episode_step[env] = ep_step; a.episode_return[env] = ep_ret; a.truncated[env] = truncated ? 1 : 0;
        if constexpr(Spec::H > 1){                         // the action-history ring goes back to its HBM rows
#pragma unroll 8
            for(int i = 0; i < 4 * Spec::H; i++) hist_hbm[(size_t)i * n] = hist_ptr[i * BLOCK];
        }
    }
    if constexpr(DR && FOLLOW){
        // Deferred parameter write-back: while the columns follow the nominal row nothing in the step loop reads the randomised entries from HBM (they live in
        // the dynamics block / the step invariants), and a reset starts from the nominal row again -- so only the LAST reset's parameters have to reach the
        // column.  They are a function of the RNG state that reset started from: re-sample once here, all lanes together, instead of 47 scattered stores
        // inside every divergent reset (~270 instructions of each reset execution).
        if(active && reset_seen){
            ParamsOverlay o;
            o.init(a.row);
            uint64_t r = rng_at_reset;
            sample_parameters<DR, Spec::RNG_OOL, B200L2F_FAST_RESET != 0>(o, r);
            o.template flush<false>(ParamsRW{a.params + env, n});
        }
    }
    if(chunk < n_chunks - 1){                             // publish: the tile's next chunk may start (on any SM)
        __threadfence();
        __syncthreads();
        if(tid == 0) atomicExch(sched + 1 + tile, chunk + 1);
    }
    }   // tile loop
    if(store_pending && lane == 0) tc::bulk_store_wait_all();   // the window must outlive the copy
    mlp_ts_epilogue(c);
}

}  // namespace b200l2f
