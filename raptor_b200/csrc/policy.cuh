// raptor_b200/csrc/policy.cuh -- actor forward passes, one environment per thread, fp32 CUDA-core path.
//
// Semantics follow the reference (INC/ = rl_tools/):
//   dense        INC/nn/layers/dense/operations_generic.h:94-108        y_o = act(b_o + sum_i W[o][i] x_i), i ascending
//   GRU step     INC/nn/layers/gru/operations_generic.h:343-411 with helper_operations_generic.h:10-71 and the generic matmul
//                INC/containers/matrix/operations_generic.h:848-868; auto-reset INC/nn/layers/gru/operations_generic.h:76-86,400-410
//   sigmoid      INC/containers/tensor/operations_generic.h:378-383     1 / (1 + exp(-a))
//   standardize  INC/nn/layers/standardize/operations_generic.h:67-84
//   squash       INC/nn/layers/sample_and_squash/operations_generic.h:148-194 (Mode<Evaluation>: tanh(mean))
//   PPO sampling INC/rl/components/on_policy_runner/operations_generic_per_env.h:43-58, log_prob INC/random/operations_generic.h:72-81
// Organisation is ours: weights are staged ONCE per launch in shared memory in k-major ("transposed") order so that for a fixed
// input index the J output weights are contiguous -> one broadcast LDS.128 feeds four FFMAs, and the J accumulators give J-way ILP.
// The summation order per output (bias first, inputs ascending) is the reference's, so results differ only by FMA contraction.
#pragma once
#include "layout.h"
#include "rng.cuh"

namespace b200l2f {

// shared-memory image of the Raptor actor (offsets in floats); K-major blocks: Wt[k * J + j] = W[j][k]
template <int IN, int HD, int OUT>
struct RaptorImage {
    static constexpr int W1T = 0;                   // [IN][HD]
    static constexpr int B1 = W1T + IN * HD;        // [HD]
    static constexpr int WIHT = B1 + HD;            // [HD][3HD]
    static constexpr int BIH = WIHT + HD * 3 * HD;  // [3HD]
    static constexpr int WHHT = BIH + 3 * HD;       // [HD][3HD]
    static constexpr int BHH = WHHT + HD * 3 * HD;  // [3HD]
    static constexpr int H0 = BHH + 3 * HD;         // [HD]
    static constexpr int W2T = H0 + HD;             // [HD][OUT]
    static constexpr int B2 = W2T + HD * OUT;       // [OUT]
    static constexpr int SIZE = B2 + OUT;
    static_assert(SIZE % 4 == 0, "image must stay float4 aligned");
};
// blob (row-major, include/b200_l2f.h) -> image (k-major); executed cooperatively by the block
template <int IN, int HD, int OUT>
__device__ __forceinline__ void stage_raptor(float* __restrict__ img, const float* __restrict__ blob){
    using I = RaptorImage<IN, HD, OUT>;
    const float* W1 = blob; const float* b1 = W1 + HD * IN;
    const float* Wih = b1 + HD; const float* bih = Wih + 3 * HD * HD;
    const float* Whh = bih + 3 * HD; const float* bhh = Whh + 3 * HD * HD;
    const float* h0 = bhh + 3 * HD; const float* W2 = h0 + HD; const float* b2 = W2 + OUT * HD;
    for(int i = threadIdx.x; i < IN * HD; i += blockDim.x){ int k = i / HD, j = i % HD; img[I::W1T + i] = W1[j * IN + k]; }
    for(int i = threadIdx.x; i < HD * 3 * HD; i += blockDim.x){ int k = i / (3 * HD), j = i % (3 * HD); img[I::WIHT + i] = Wih[j * HD + k]; img[I::WHHT + i] = Whh[j * HD + k]; }
    for(int i = threadIdx.x; i < HD * OUT; i += blockDim.x){ int k = i / OUT, j = i % OUT; img[I::W2T + i] = W2[j * HD + k]; }
    for(int i = threadIdx.x; i < HD; i += blockDim.x){ img[I::B1 + i] = b1[i]; img[I::H0 + i] = h0[i]; }
    for(int i = threadIdx.x; i < 3 * HD; i += blockDim.x){ img[I::BIH + i] = bih[i]; img[I::BHH + i] = bhh[i]; }
    for(int i = threadIdx.x; i < OUT; i += blockDim.x) img[I::B2 + i] = b2[i];
}

// Two homes for the actor image:
//   WeightsShared  staged in shared memory, read with broadcast LDS.128 (one load feeds four FFMAs)
//   WeightsParam   passed BY VALUE as a __grid_constant__ kernel parameter: the weights live in the constant bank of the launch and
//                  every FFMA takes its weight as a c[0x0][imm] operand -- no load instruction, no register, no scoreboard wait.
struct WeightsShared {
    const float* w;
    __device__ __forceinline__ float operator[](int i) const { return w[i]; }
    __device__ __forceinline__ float4 load4(int i) const { return *reinterpret_cast<const float4*>(w + i); }
};
template <int SIZE>
struct WeightBlock { float w[SIZE]; };
template <int SIZE>
struct WeightsParam {
    const WeightBlock<SIZE>& b;
    __device__ __forceinline__ float operator[](int i) const { return b.w[i]; }
    __device__ __forceinline__ float4 load4(int i) const { return make_float4(b.w[i], b.w[i + 1], b.w[i + 2], b.w[i + 3]); }
};

// host-side twin of stage_raptor (same k-major image), used to build the by-value kernel parameter
template <int IN, int HD, int OUT>
inline void build_raptor_image_host(float* img, const float* blob){
    using I = RaptorImage<IN, HD, OUT>;
    const float* W1 = blob; const float* b1 = W1 + HD * IN;
    const float* Wih = b1 + HD; const float* bih = Wih + 3 * HD * HD;
    const float* Whh = bih + 3 * HD; const float* bhh = Whh + 3 * HD * HD;
    const float* h0 = bhh + 3 * HD; const float* W2 = h0 + HD; const float* b2 = W2 + OUT * HD;
    for(int k = 0; k < IN; k++) for(int j = 0; j < HD; j++) img[I::W1T + k * HD + j] = W1[j * IN + k];
    for(int k = 0; k < HD; k++) for(int j = 0; j < 3 * HD; j++){ img[I::WIHT + k * 3 * HD + j] = Wih[j * HD + k]; img[I::WHHT + k * 3 * HD + j] = Whh[j * HD + k]; }
    for(int k = 0; k < HD; k++) for(int j = 0; j < OUT; j++) img[I::W2T + k * OUT + j] = W2[j * HD + k];
    for(int i = 0; i < HD; i++){ img[I::B1 + i] = b1[i]; img[I::H0 + i] = h0[i]; }
    for(int i = 0; i < 3 * HD; i++){ img[I::BIH + i] = bih[i]; img[I::BHH + i] = bhh[i]; }
    for(int i = 0; i < OUT; i++) img[I::B2 + i] = b2[i];
}

// acc[j] += Wt[k][j] * x[k], k ascending (outer), j inner: J independent FFMA chains
template <int J, int K, class W>
__device__ __forceinline__ void matvec_acc(float* __restrict__ acc, const W& img, int wt, const float* __restrict__ x){
    static_assert(J % 4 == 0, "J must be a multiple of 4");
#pragma unroll
    for(int k = 0; k < K; k++){
        const float xk = x[k];
#pragma unroll
        for(int j4 = 0; j4 < J / 4; j4++){
            const float4 w = img.load4(wt + k * J + 4 * j4);
            acc[4 * j4 + 0] += w.x * xk;
            acc[4 * j4 + 1] += w.y * xk;
            acc[4 * j4 + 2] += w.z * xk;
            acc[4 * j4 + 3] += w.w * xk;
        }
    }
}

template <bool FAST>
__device__ __forceinline__ float sigmoidf_(float a){
    if constexpr(FAST) return sigmoid_of_scaled(a * -LOG2E);   // 1 / (1 + e^-a); saturates to 0 / 1 through ex2 -> inf / 0
    else return 1.0f / (1.0f + expf(-a));
}
template <bool FAST>
__device__ __forceinline__ float tanhf_(float a){
    if constexpr(FAST) return tanh_of_scaled(a * (2.0f * LOG2E));   // 1 - 2 / (e^{2a} + 1): absolute error ~1e-7, saturates correctly
    else return tanhf(a);
}

// Raptor actor: Dense(IN->HD, ReLU) -> GRU(HD) -> Dense(HD->OUT).  h (registers) and gru_step are updated in place.
template <int IN, int HD, int OUT, bool FAST, class W>
__device__ __forceinline__ void raptor_forward(const W& img, const float* __restrict__ obs, float* __restrict__ h, int& gru_step,
                                               int seq_len, bool no_auto_reset, float* __restrict__ action){
    using I = RaptorImage<IN, HD, OUT>;
    float x1[HD];
#pragma unroll
    for(int j = 0; j < HD; j++) x1[j] = img[I::B1 + j];
    matvec_acc<HD, IN>(x1, img, I::W1T, obs);
#pragma unroll
    for(int j = 0; j < HD; j++) x1[j] = fmaxf(x1[j], 0.0f);
    // reset_truncate (gru/operations_generic.h:76-86)
    if(!no_auto_reset && gru_step >= seq_len){
#pragma unroll
        for(int j = 0; j < HD; j++) h[j] = img[I::H0 + j];
        gru_step = 0;
    }
    float r[HD], z[HD];
    {   // r, z = sigmoid((b_hh + W_hh h) + b_ih + W_ih x)
        float pre[2 * HD];
#pragma unroll
        for(int j = 0; j < 2 * HD; j++) pre[j] = img[I::BHH + j];
#pragma unroll
        for(int k = 0; k < HD; k++){
            const float hk = h[k];
#pragma unroll
            for(int j4 = 0; j4 < 2 * HD / 4; j4++){
                const float4 w = img.load4(I::WHHT + k * 3 * HD + 4 * j4);
                pre[4 * j4 + 0] += w.x * hk; pre[4 * j4 + 1] += w.y * hk; pre[4 * j4 + 2] += w.z * hk; pre[4 * j4 + 3] += w.w * hk;
            }
        }
#pragma unroll
        for(int j = 0; j < 2 * HD; j++) pre[j] += img[I::BIH + j];
#pragma unroll
        for(int k = 0; k < HD; k++){
            const float xk = x1[k];
#pragma unroll
            for(int j4 = 0; j4 < 2 * HD / 4; j4++){
                const float4 w = img.load4(I::WIHT + k * 3 * HD + 4 * j4);
                pre[4 * j4 + 0] += w.x * xk; pre[4 * j4 + 1] += w.y * xk; pre[4 * j4 + 2] += w.z * xk; pre[4 * j4 + 3] += w.w * xk;
            }
        }
#pragma unroll
        for(int j = 0; j < HD; j++){ r[j] = sigmoidf_<FAST>(pre[j]); z[j] = sigmoidf_<FAST>(pre[HD + j]); }
    }
    float hn[HD];
    {   // n = tanh((b_in + W_in x) + (b_hn + W_hn h) * r);  h' = (1 - z) * n + z * h
        float nh[HD], nx[HD];
#pragma unroll
        for(int j = 0; j < HD; j++){ nh[j] = img[I::BHH + 2 * HD + j]; nx[j] = img[I::BIH + 2 * HD + j]; }
#pragma unroll
        for(int k = 0; k < HD; k++){
            const float hk = h[k], xk = x1[k];
#pragma unroll
            for(int j4 = 0; j4 < HD / 4; j4++){
                const float4 wh = img.load4(I::WHHT + k * 3 * HD + 2 * HD + 4 * j4);
                const float4 wx = img.load4(I::WIHT + k * 3 * HD + 2 * HD + 4 * j4);
                nh[4 * j4 + 0] += wh.x * hk; nh[4 * j4 + 1] += wh.y * hk; nh[4 * j4 + 2] += wh.z * hk; nh[4 * j4 + 3] += wh.w * hk;
                nx[4 * j4 + 0] += wx.x * xk; nx[4 * j4 + 1] += wx.y * xk; nx[4 * j4 + 2] += wx.z * xk; nx[4 * j4 + 3] += wx.w * xk;
            }
        }
#pragma unroll
        for(int j = 0; j < HD; j++){
            const float n = tanhf_<FAST>(nx[j] + nh[j] * r[j]);
            hn[j] = (1.0f - z[j]) * n + z[j] * h[j];
        }
    }
    {
        float a[OUT];
#pragma unroll
        for(int j = 0; j < OUT; j++) a[j] = img[I::B2 + j];
        matvec_acc<OUT, HD>(a, img, I::W2T, hn);
#pragma unroll
        for(int j = 0; j < OUT; j++) action[j] = a[j];
    }
    int new_step = gru_step + 1;
    const bool wrap = !no_auto_reset && new_step >= seq_len;   // gru/operations_generic.h:400-410: the OUTPUT of this step is kept, the stored state resets
#pragma unroll
    for(int j = 0; j < HD; j++) h[j] = wrap ? img[I::H0 + j] : hn[j];
    gru_step = wrap ? 0 : new_step;
}

// ---------------------------------------------------------------------------------------------------------------
// Compact-code variant used by the fused kernels: the k loops are ROLLED (the fully unrolled actor is ~2500 instructions = 40 KB of
// SASS per copy and, together with four inlined RK4 stages, overflows the instruction cache: ncu shows sm__icc_request_hit_rate 72-88 %
// and a "no_instruction" stall).  Vectors that are indexed by the loop variable (observation, dense-1 output, hidden state) live in a
// per-thread shared-memory column scr[row * stride] -- the hidden state stays there for the whole launch -- and the J accumulators of
// a layer stay in registers.  Summation order per output is unchanged (bias, then inputs ascending; W_hh*h before b_ih + W_ih*x).
// scratch rows: [0, IN) observation, [IN, IN+HD) dense-1 output, [IN+HD, IN+2HD) hidden state
// ---------------------------------------------------------------------------------------------------------------
template <int IN, int HD>
struct RaptorScratch {
    static constexpr int OBS = 0, X1 = IN, H = IN + HD, ROWS = IN + 2 * HD;
};
template <int IN, int HD, int OUT, bool FAST>
__device__ __forceinline__ void raptor_forward_rolled(const float* __restrict__ img, float* __restrict__ scr, int stride, int& gru_step,
                                                      int seq_len, bool no_auto_reset, float* __restrict__ action){
    using I = RaptorImage<IN, HD, OUT>;
    using S = RaptorScratch<IN, HD>;
    {   // dense 1 + ReLU -> scratch
        float x1[HD];
#pragma unroll
        for(int j = 0; j < HD; j++) x1[j] = img[I::B1 + j];
#pragma unroll 2
        for(int k = 0; k < IN; k++){
            const float xk = scr[(S::OBS + k) * stride];
            const float* w = img + I::W1T + k * HD;
#pragma unroll
            for(int j4 = 0; j4 < HD / 4; j4++){
                const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
                x1[4 * j4 + 0] += w4.x * xk; x1[4 * j4 + 1] += w4.y * xk; x1[4 * j4 + 2] += w4.z * xk; x1[4 * j4 + 3] += w4.w * xk;
            }
        }
#pragma unroll
        for(int j = 0; j < HD; j++) scr[(S::X1 + j) * stride] = fmaxf(x1[j], 0.0f);
    }
    if(!no_auto_reset && gru_step >= seq_len){   // reset_truncate (gru/operations_generic.h:76-86)
#pragma unroll
        for(int j = 0; j < HD; j++) scr[(S::H + j) * stride] = img[I::H0 + j];
        gru_step = 0;
    }
    float pre[3 * HD], nx[HD];
#pragma unroll
    for(int j = 0; j < 3 * HD; j++) pre[j] = img[I::BHH + j];
#pragma unroll 2
    for(int k = 0; k < HD; k++){   // pre = b_hh + W_hh h
        const float hk = scr[(S::H + k) * stride];
        const float* w = img + I::WHHT + k * 3 * HD;
#pragma unroll
        for(int j4 = 0; j4 < 3 * HD / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
            pre[4 * j4 + 0] += w4.x * hk; pre[4 * j4 + 1] += w4.y * hk; pre[4 * j4 + 2] += w4.z * hk; pre[4 * j4 + 3] += w4.w * hk;
        }
    }
#pragma unroll
    for(int j = 0; j < 2 * HD; j++) pre[j] += img[I::BIH + j];
#pragma unroll
    for(int j = 0; j < HD; j++) nx[j] = img[I::BIH + 2 * HD + j];
#pragma unroll 2
    for(int k = 0; k < HD; k++){   // r,z pre-activations += W_ih x ;  n_x = b_in + W_in x
        const float xk = scr[(S::X1 + k) * stride];
        const float* w = img + I::WIHT + k * 3 * HD;
#pragma unroll
        for(int j4 = 0; j4 < 2 * HD / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
            pre[4 * j4 + 0] += w4.x * xk; pre[4 * j4 + 1] += w4.y * xk; pre[4 * j4 + 2] += w4.z * xk; pre[4 * j4 + 3] += w4.w * xk;
        }
#pragma unroll
        for(int j4 = 0; j4 < HD / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 2 * HD + 4 * j4);
            nx[4 * j4 + 0] += w4.x * xk; nx[4 * j4 + 1] += w4.y * xk; nx[4 * j4 + 2] += w4.z * xk; nx[4 * j4 + 3] += w4.w * xk;
        }
    }
    float a[OUT];
#pragma unroll
    for(int j = 0; j < OUT; j++) a[j] = img[I::B2 + j];
    const bool wrap = !no_auto_reset && gru_step + 1 >= seq_len;   // gru/operations_generic.h:400-410: this step's OUTPUT is kept, the stored state resets
#pragma unroll
    for(int j = 0; j < HD; j++){
        const float r = sigmoidf_<FAST>(pre[j]);
        const float z = sigmoidf_<FAST>(pre[HD + j]);
        const float n = tanhf_<FAST>(nx[j] + pre[2 * HD + j] * r);
        const float hn = (1.0f - z) * n + z * scr[(S::H + j) * stride];
        const float4 w4 = *reinterpret_cast<const float4*>(img + I::W2T + j * OUT);   // OUT == 4
        a[0] += w4.x * hn; a[1] += w4.y * hn; a[2] += w4.z * hn; a[3] += w4.w * hn;
        scr[(S::H + j) * stride] = wrap ? img[I::H0 + j] : hn;
    }
    gru_step = wrap ? 0 : gru_step + 1;
#pragma unroll
    for(int j = 0; j < OUT; j++) action[j] = a[j];
}

// ---------------------------------------------------------------------------------------------------------------
// MLP actors (SAC teacher / PPO): [standardize] -> Dense(IN->HD, ReLU) -> Dense(HD->HD, ReLU) -> Dense(HD->OUT)
// image: [mean[IN] precision[IN]] W1T[IN][HD] b1[HD] W2T[HD][HD] b2[HD] W3T[HD][OUT] b3[OUT] [log_std[4]]
// ---------------------------------------------------------------------------------------------------------------
template <int IN, int HD, int OUT, bool STD, bool LOGSTD>
struct MlpImage {
    static constexpr int MEAN = 0;
    static constexpr int PREC = MEAN + (STD ? IN : 0);
    static constexpr int W1T = PREC + (STD ? IN : 0);
    static constexpr int B1 = W1T + IN * HD;
    static constexpr int W2T = B1 + HD;
    static constexpr int B2 = W2T + HD * HD;
    static constexpr int W3T = B2 + HD;
    static constexpr int B3 = W3T + HD * OUT;
    static constexpr int LOG_STD = B3 + OUT;
    static constexpr int SIZE_RAW = LOG_STD + (LOGSTD ? 4 : 0);
    static constexpr int SIZE = (SIZE_RAW + 3) / 4 * 4;
    static_assert((!STD || IN % 2 == 0) && (IN * HD) % 4 == 0, "float4 alignment of the weight blocks");
};
template <int IN, int HD, int OUT, bool STD, bool LOGSTD>
__device__ __forceinline__ void stage_mlp(float* __restrict__ img, const float* __restrict__ blob){
    using I = MlpImage<IN, HD, OUT, STD, LOGSTD>;
    const float* b = blob;
    if constexpr(STD){
        for(int i = threadIdx.x; i < 2 * IN; i += blockDim.x) img[I::MEAN + i] = b[i];
        b += 2 * IN;
    }
    const float* W1 = b; const float* b1 = W1 + HD * IN; const float* W2 = b1 + HD; const float* b2 = W2 + HD * HD;
    const float* W3 = b2 + HD; const float* b3 = W3 + OUT * HD; const float* ls = b3 + OUT;
    for(int i = threadIdx.x; i < IN * HD; i += blockDim.x){ int k = i / HD, j = i % HD; img[I::W1T + i] = W1[j * IN + k]; }
    for(int i = threadIdx.x; i < HD * HD; i += blockDim.x){ int k = i / HD, j = i % HD; img[I::W2T + i] = W2[j * HD + k]; }
    for(int i = threadIdx.x; i < HD * OUT; i += blockDim.x){ int k = i / OUT, j = i % OUT; img[I::W3T + i] = W3[j * HD + k]; }
    for(int i = threadIdx.x; i < HD; i += blockDim.x){ img[I::B1 + i] = b1[i]; img[I::B2 + i] = b2[i]; }
    for(int i = threadIdx.x; i < OUT; i += blockDim.x) img[I::B3 + i] = b3[i];
    if constexpr(LOGSTD){ for(int i = threadIdx.x; i < 4; i += blockDim.x) img[I::LOG_STD + i] = ls[i]; }
}

}  // namespace b200l2f
