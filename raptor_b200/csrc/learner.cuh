// raptor_b200/csrc/learner.cuh -- the learner feed of the PPO loop step, i.e. what the reference does between rl_tools::collect and
// rl_tools::train on the dataset the collection kernel wrote (INC/rl/algorithms/ppo/loop/core/operations_generic.h:104-117):
//
//   k_values_ts<IN, GAE>   critic (standardize -> Dense 64 ReLU -> Dense 64 ReLU -> Dense 1) over all (T+1) n observation rows -> the all_values
//                          column (`evaluate(device, ts.ppo.critic, all_observations_privileged, all_values, ...)`, :112-116) and, with GAE, the
//                          generalized advantage estimation (`estimate_generalized_advantages`, INC/rl/algorithms/ppo/operations_generic.h:54-89)
//                          in the SAME backward pass over time: the dataset is read once.
//                          One CTA = 128 environments; per step the four warps pull their 32 consecutive dataset rows (32 x D floats, contiguous)
//                          into shared memory with one TMA bulk copy each, double-buffered, so the copy of step t-1 runs under the GEMMs of
//                          step t.  The three GEMMs are the tcgen05 / TMEM path of mlp_tc.cuh (3xTF32, fp32-equivalent accuracy).
//   k_values<IN>           the same critic on fp32 CUDA cores (B200L2F_GEMM_FP32_CUDA_CORES / B200L2F_FLAG_ACCURATE_MATH), one row per thread
//   k_gae                  stand-alone GAE for callers that computed the values themselves: one environment per thread, touches only the
//                          six columns it needs
//   k_column_partials / k_column_finish   column mean and sample standard deviation of the observation block for the running observation
//                          normalizer (INC/rl/components/running_normalizer/operations_generic.h:27-49): deterministic two-level reduction
//
// Dataset layout: rows = step * n + env, D = OBS + 15 columns (on_policy_runner.h:42-64, see mlp.cuh above CollectArgs).
#pragma once
#include "mlp_tc.cuh"

namespace b200l2f {

struct FeedArgs {
    float* dataset;        // [(T+1) n][D]
    int n, T;
    float gamma, lambda;
    int ignore_termination;
    const float* blob;     // critic blob, output padded to 4 rows (CUDA-core kernel)
    int has_std;
    int* sched;            // persistent tile counter
};

// reference GAE recursion for one (environment, step), going backwards in time
struct GaeCarry { float previous_value, previous_advantage; };
__device__ __forceinline__ void gae_step(GaeCarry& g, float reward, bool terminated, bool truncated, float value, float gamma, float lambda, bool ignore_termination,
                                         float& advantage, float& target_value){
    const bool terminated_actual = terminated && !ignore_termination;
    const float next_step_value = terminated_actual ? 0.0f : g.previous_value;
    float td_error = B200_SUB(B200_ADD(reward, B200_MUL(gamma, next_step_value)), value);
    if(truncated){
        if(!terminated) td_error = 0.0f;      // time limit / random truncation
        g.previous_advantage = 0.0f;
    }
    advantage = B200_ADD(B200_MUL(B200_MUL(lambda, gamma), g.previous_advantage), td_error);
    target_value = B200_ADD(advantage, value);
    g.previous_advantage = advantage;
    g.previous_value = value;
}

template <int IN>
struct FeedSmem {
    static constexpr int D = IN + 15;
    static constexpr int WIN_FLOATS = 32 * D;                             // one warp's 32 rows
    static constexpr int WIN_BYTES = WIN_FLOATS * 4;
    static_assert(WIN_BYTES % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
    static constexpr int B = 0;
    static constexpr int BAR = (MlpTcImage<IN, 4>::BYTES + 127) / 128 * 128;   // bar_tma, bar_mma, tmem slot | 8 row barriers from +32
    static constexpr int ROWS = BAR + 128;
    static constexpr int TOTAL = ROWS + 4 * 2 * WIN_BYTES;
};

template <int IN, bool GAE>
__global__ void __launch_bounds__(BLOCK, 2) k_values_ts(const __grid_constant__ FeedArgs a, const float* __restrict__ tc_image){
    constexpr int D = IN + 15;
    using SM = FeedSmem<IN>;
    extern __shared__ __align__(1024) unsigned char smraw[];
    TsCtx c = mlp_ts_prologue_at<IN, 4>(smraw, SM::B, SM::BAR, tc_image);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint64_t* row_bar = reinterpret_cast<uint64_t*>(smraw + SM::BAR + 32) + warp * 2;
    float* win = reinterpret_cast<float*>(smraw + SM::ROWS) + (size_t)warp * 2 * SM::WIN_FLOATS;
    if(lane == 0){ tc::mbar_init(row_bar, 1); tc::mbar_init(row_bar + 1, 1); tc::mbar_fence_init(); }
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    const size_t n = (size_t)a.n;
    // bulk copies need 16-byte aligned sources: the dataset base and the row pitch of one step (n * D * 4 bytes)
    const bool aligned = ((reinterpret_cast<uintptr_t>(a.dataset) & 15) == 0) && ((n * D) % 4 == 0);
    __shared__ int s_item;
    const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
    for(;;){
        if(tid == 0) s_item = atomicAdd(a.sched, 1);
        __syncthreads();
        const int tile = s_item;
        __syncthreads();
        if(tile >= n_tiles) break;
        const int warp_env0 = tile * BLOCK + warp * 32;
        const int rows_valid = min(32, a.n - warp_env0);     // <= 0: this warp only takes part in the CTA-collective GEMMs
        const bool active = lane < rows_valid;
        const bool bulk = aligned && rows_valid == 32;
        const size_t env = (size_t)warp_env0 + lane;
        auto fetch = [&](int t, int buf){
            if(rows_valid <= 0) return;
            const float* src = a.dataset + ((size_t)t * n + warp_env0) * D;
            float* dst = win + buf * SM::WIN_FLOATS;
            if(bulk){
                if(tc::elect_one()){                           // all 32 lanes arrive here together (warp-uniform branch after __syncwarp)
                    tc::fence_async_smem();                     // earlier generic-proxy reads of this window are ordered before the async-proxy write
                    tc::mbar_expect_tx(row_bar + buf, SM::WIN_BYTES);
                    tc::tma_load_1d(dst, src, SM::WIN_BYTES, row_bar + buf);
                }
            }
            else{
                const int count = rows_valid * D;
                for(int idx = lane; idx < count; idx += 32) dst[idx] = src[idx];
            }
        };
        fetch(a.T, 0);
        GaeCarry g{0.0f, 0.0f};
        for(int t = a.T; t >= 0; t--){
            const int buf = (a.T - t) & 1;
            __syncwarp();                                       // every lane is done with the other window (read two steps ago)
            if(t > 0) fetch(t - 1, buf ^ 1);
            if(bulk){
                if(buf == 0){ tc::mbar_wait(row_bar, ph0); ph0 ^= 1; }
                else{ tc::mbar_wait(row_bar + 1, ph1); ph1 ^= 1; }
            }
            __syncwarp();
            const float* row = win + buf * SM::WIN_FLOATS + lane * D;   // D is odd: conflict-free
            float obs[IN];
#pragma unroll
            for(int i = 0; i < IN; i++) obs[i] = active ? row[i] : 0.0f;
            float reward = 0.0f; bool terminated = false, truncated = false;
            if(GAE && active && t < a.T){ reward = row[IN + 9]; terminated = row[IN + 10] != 0.0f; truncated = row[IN + 11] != 0.0f; }
            float v[4];
            mlp_forward_ts<IN, 4>(c, obs, v);
            if(active){
                float* out = a.dataset + ((size_t)t * n + env) * D + IN + 12;
                out[0] = v[0];
                if constexpr(GAE){
                    if(t == a.T) g.previous_value = v[0];
                    else{
                        float adv, target;
                        gae_step(g, reward, terminated, truncated, v[0], a.gamma, a.lambda, a.ignore_termination != 0, adv, target);
                        out[1] = adv; out[2] = target;
                    }
                }
            }
        }
    }
    mlp_ts_epilogue(c);
}

// fp32 CUDA-core critic: one dataset row per thread
template <int IN>
__global__ void __launch_bounds__(BLOCK) k_values(const __grid_constant__ FeedArgs a){
    constexpr int D = IN + 15;
    extern __shared__ __align__(16) float smem[];
    float* img = smem;
    float* scr = smem + MlpImg<IN, 4>::SIZE + threadIdx.x;
    stage_mlp_image<IN, 4>(img, a.blob, a.has_std != 0, false);
    __syncthreads();
    const size_t rows = (size_t)(a.T + 1) * a.n;
    for(size_t r = (size_t)blockIdx.x * BLOCK + threadIdx.x; r < rows; r += (size_t)gridDim.x * BLOCK){
        float* row = a.dataset + r * D;
        for(int i = 0; i < IN; i++) scr[i * BLOCK] = row[i];
        float o[4];
        mlp_forward<IN, 4>(img, scr, BLOCK, o);
        row[IN + 12] = o[0];
    }
}

// stand-alone GAE: one environment per thread, backwards in time
__global__ void k_gae(const __grid_constant__ FeedArgs a, int obs_dim){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= a.n) return;
    const int D = obs_dim + 15;
    const size_t n = (size_t)a.n;
    GaeCarry g{a.dataset[((size_t)a.T * n + e) * D + obs_dim + 12], 0.0f};
    for(int t = a.T - 1; t >= 0; t--){
        float* row = a.dataset + ((size_t)t * n + e) * D + obs_dim + 9;    // reward terminated truncated value advantage target_value
        float adv, target;
        gae_step(g, row[0], row[1] != 0.0f, row[2] != 0.0f, row[3], a.gamma, a.lambda, a.ignore_termination != 0, adv, target);
        row[4] = adv; row[5] = target;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// column statistics of the observation block (rows [0, T n), columns [0, OBS)): pass 0 sums x, pass 1 sums (x - mean)^2.
// Each CTA stages 128 consecutive rows (contiguous in memory, coalesced) and thread c < OBS adds column c; per-CTA partials in double,
// reduced in a fixed order by k_column_finish: deterministic, and more accurate than the reference's sequential fp32 sums.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_column_partials(const float* __restrict__ dataset, size_t rows, int obs_dim, const double* __restrict__ mean, double* __restrict__ partials){
    extern __shared__ __align__(16) float tile[];     // [BLOCK][D]
    const int D = obs_dim + 15;
    double acc = 0.0;
    const double mu = (mean && threadIdx.x < obs_dim) ? mean[threadIdx.x] : 0.0;
    for(size_t r0 = (size_t)blockIdx.x * BLOCK; r0 < rows; r0 += (size_t)gridDim.x * BLOCK){
        const int nr = (int)min((size_t)BLOCK, rows - r0);
        const float* src = dataset + r0 * D;
        __syncthreads();
        for(int idx = threadIdx.x; idx < nr * D; idx += BLOCK) tile[idx] = src[idx];
        __syncthreads();
        if(threadIdx.x < obs_dim){
            if(mean){ for(int r = 0; r < nr; r++){ const double d = (double)tile[r * D + threadIdx.x] - mu; acc += d * d; } }
            else{ for(int r = 0; r < nr; r++) acc += (double)tile[r * D + threadIdx.x]; }
        }
    }
    if(threadIdx.x < obs_dim) partials[(size_t)blockIdx.x * obs_dim + threadIdx.x] = acc;
}
// pass 0: out[c] = sum / rows.  pass 1: out[c] = sum (the caller applies the reference's `acc < 1e-6 -> 0` rule and the sqrt)
__global__ void k_column_finish(const double* __restrict__ partials, int n_partials, int obs_dim, double scale, double* __restrict__ out){
    const int c = threadIdx.x;
    if(c >= obs_dim) return;
    double acc = 0.0;
    for(int b = 0; b < n_partials; b++) acc += partials[(size_t)b * obs_dim + c];
    out[c] = acc * scale;
}

}  // namespace b200l2f
