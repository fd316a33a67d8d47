// raptor_b200/csrc/mlp.cuh -- MLP actors (SAC teacher / PPO) and the two loop owners built on them:
//   k_rollout_mlp   closed-loop rollout with a deterministic MLP actor (rl_tools::evaluate order), BASELINE config 3
//   k_collect       PPO collection with on-device auto-reset, Gaussian action sampling and trajectory write-back (rl_tools::collect),
//                   BASELINE config 4
// Reference semantics (INC/ = rl_tools/):
//   MLP            INC/nn_models/mlp/network.h:15-51 (input -> HD ReLU -> HD ReLU -> OUT identity), dense INC/nn/layers/dense/operations_generic.h:94-108
//   standardize    INC/nn/layers/standardize/operations_generic.h:67-84
//   squash (eval)  INC/nn/layers/sample_and_squash/operations_generic.h:148-194 (Mode<Evaluation>: tanh(mean), no draw)
//   PPO sampling   INC/rl/components/on_policy_runner/operations_generic_per_env.h:43-58, log_prob INC/random/operations_generic.h:72-81
//   collect        INC/rl/components/on_policy_runner/operations_generic.h:99-131 + operations_generic_per_env.h:8-75; dataset columns
//                  INC/rl/components/on_policy_runner/on_policy_runner.h:42-64, operations_generic.h:12-29
// Organisation: one environment per thread; the actor's weights are staged once per CTA in shared memory in k-major order; the vector a
// layer consumes lives in a per-thread shared-memory column (so the k loops can stay rolled and the 64 accumulators of a layer are the
// only wide register array); dataset rows are transposed through shared memory and written as fully coalesced 128-byte lines.
#pragma once
#include <type_traits>
#include "kernels.cuh"

namespace b200l2f {

constexpr int MLP_HD = 64;

// shared-memory image (floats): mean[IN] precision[IN] W1T[IN][HD] b1[HD] W2T[HD][HD] b2[HD] W3T[HD][OUT] b3[OUT] log_std[4]
template <int IN, int OUT>
struct MlpImg {
    static constexpr int HD = MLP_HD;
    static constexpr int IN_PAD = (IN + 3) / 4 * 4;
    static constexpr int MEAN = 0, PREC = MEAN + IN_PAD, W1T = PREC + IN_PAD, B1 = W1T + IN * HD, W2T = B1 + HD, B2 = W2T + HD * HD,
                         W3T = B2 + HD, B3 = W3T + HD * OUT, LOG_STD = B3 + OUT, SIZE = LOG_STD + 4;
    static_assert(SIZE % 4 == 0 && W1T % 4 == 0 && W3T % 4 == 0, "float4 alignment");
};
// blob (include/b200_l2f.h MLP order; standardize / log_std blocks optional) -> image; missing standardize = (mean 0, precision 1)
template <int IN, int OUT>
__device__ __forceinline__ void stage_mlp_image(float* __restrict__ img, const float* __restrict__ blob, bool has_std, bool has_log_std){
    using I = MlpImg<IN, OUT>;
    constexpr int HD = MLP_HD;
    const float* b = blob;
    for(int i = threadIdx.x; i < IN; i += blockDim.x){ img[I::MEAN + i] = has_std ? b[i] : 0.0f; img[I::PREC + i] = has_std ? b[IN + i] : 1.0f; }
    if(has_std) b += 2 * IN;
    const float* W1 = b; const float* b1 = W1 + HD * IN; const float* W2 = b1 + HD; const float* b2 = W2 + HD * HD;
    const float* W3 = b2 + HD; const float* b3 = W3 + OUT * HD; const float* ls = b3 + OUT;
    for(int i = threadIdx.x; i < IN * HD; i += blockDim.x){ int k = i / HD, j = i % HD; img[I::W1T + i] = W1[j * IN + k]; }
    for(int i = threadIdx.x; i < HD * HD; i += blockDim.x){ int k = i / HD, j = i % HD; img[I::W2T + i] = W2[j * HD + k]; }
    for(int i = threadIdx.x; i < HD * OUT; i += blockDim.x){ int k = i / OUT, j = i % OUT; img[I::W3T + i] = W3[j * HD + k]; }
    for(int i = threadIdx.x; i < HD; i += blockDim.x){ img[I::B1 + i] = b1[i]; img[I::B2 + i] = b2[i]; }
    for(int i = threadIdx.x; i < OUT; i += blockDim.x) img[I::B3 + i] = b3[i];
    for(int i = threadIdx.x; i < 4; i += blockDim.x) img[I::LOG_STD + i] = has_log_std ? ls[i] : 0.0f;
}

// scr: this thread's scratch column (rows scr[r * stride]); rows [IN_ROW0, IN_ROW0 + IN) hold the RAW observation on entry, rows [0, 64) are
// overwritten with the hidden activations (IN_ROW0 = 0: the observation is consumed before it is overwritten; IN_ROW0 = 64: it survives).
template <int IN, int OUT, int IN_ROW0 = 0>
__device__ __forceinline__ void mlp_forward(const float* __restrict__ img, float* __restrict__ scr, int stride, float* __restrict__ out){
    using I = MlpImg<IN, OUT>;
    constexpr int HD = MLP_HD;
    float acc[HD];
#pragma unroll
    for(int j = 0; j < HD; j++) acc[j] = img[I::B1 + j];
#pragma unroll 2
    for(int k = 0; k < IN; k++){
        float xk = scr[(IN_ROW0 + k) * stride] - img[I::MEAN + k];          // standardize: (x - mean) [* precision unless it is 0]
        const float pr = img[I::PREC + k];
        if(pr != 0.0f) xk *= pr;
        const float* w = img + I::W1T + k * HD;
#pragma unroll
        for(int j4 = 0; j4 < HD / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
            acc[4 * j4] += w4.x * xk; acc[4 * j4 + 1] += w4.y * xk; acc[4 * j4 + 2] += w4.z * xk; acc[4 * j4 + 3] += w4.w * xk;
        }
    }
#pragma unroll
    for(int j = 0; j < HD; j++){ scr[j * stride] = fmaxf(acc[j], 0.0f); acc[j] = img[I::B2 + j]; }
#pragma unroll 2
    for(int k = 0; k < HD; k++){
        const float xk = scr[k * stride];
        const float* w = img + I::W2T + k * HD;
#pragma unroll
        for(int j4 = 0; j4 < HD / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
            acc[4 * j4] += w4.x * xk; acc[4 * j4 + 1] += w4.y * xk; acc[4 * j4 + 2] += w4.z * xk; acc[4 * j4 + 3] += w4.w * xk;
        }
    }
#pragma unroll
    for(int j = 0; j < HD; j++) scr[j * stride] = fmaxf(acc[j], 0.0f);
    float o[OUT];
#pragma unroll
    for(int j = 0; j < OUT; j++) o[j] = img[I::B3 + j];
#pragma unroll 4
    for(int k = 0; k < HD; k++){
        const float xk = scr[k * stride];
        const float* w = img + I::W3T + k * OUT;
#pragma unroll
        for(int j4 = 0; j4 < OUT / 4; j4++){
            const float4 w4 = *reinterpret_cast<const float4*>(w + 4 * j4);
            o[4 * j4] += w4.x * xk; o[4 * j4 + 1] += w4.y * xk; o[4 * j4 + 2] += w4.z * xk; o[4 * j4 + 3] += w4.w * xk;
        }
    }
#pragma unroll
    for(int j = 0; j < OUT; j++) out[j] = o[j];
}

// full observation of the spec into the scratch column (and optionally a global row)
template <class Spec, bool FAST = false, class P>
__device__ __forceinline__ void observe_to_scratch(const EnvState<Spec>& st, const P& p, uint64_t& rng, const float* hist_ptr, size_t n,
                                                   float* __restrict__ scr, int stride){
    float o[18];
    observe18<Spec, true, FAST>(st, p, rng, o);
#pragma unroll
    for(int i = 0; i < 18; i++) scr[i * stride] = o[i];
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) scr[(18 + i) * stride] = st.hist[i];
    }
    else{
        int cur = st.current_step == 0 ? Spec::H - 1 : st.current_step - 1;
        for(int h = 0; h < Spec::H; h++){
#pragma unroll
            for(int i = 0; i < 4; i++) scr[(18 + 4 * h + i) * stride] = hist_ptr[(size_t)(4 * cur + i) * n];
            cur = cur == 0 ? Spec::H - 1 : cur - 1;
        }
    }
    if constexpr(Spec::OBS_LAYOUT == OBS_TEACHER){
#pragma unroll
        for(int i = 0; i < 4; i++) scr[(18 + 4 * Spec::H + i) * stride] = (st.x[X_RPM + i] - p[P_ACT_MIN]) / (p[P_ACT_MAX] - p[P_ACT_MIN]) * 2.0f - 1.0f;
    }
}

// action head.  SQUASH (OUT == 8): tanh(mean) of the first four outputs.  PPO: a ~ N(mean, exp(log_std)), summed log-probability.
__device__ __forceinline__ float normal_log_prob(float mean, float log_std, float value){
    const float neg_log_sqrt_pi = -0.5f * logf(2.0f * 3.14159274101257324f);
    const float pre = (value - mean) / expf(log_std);
    return neg_log_sqrt_pi - log_std - 0.5f * pre * pre;
}

// ---------------------------------------------------------------------------------------------------------------
// vector-API kernel: one evaluate_step of an MLP actor on caller-provided observations
// ---------------------------------------------------------------------------------------------------------------
template <int IN, int OUT>
__global__ void __launch_bounds__(BLOCK) k_mlp_step(const float* __restrict__ blob, int has_std, int has_log_std, int head, const float* __restrict__ obs, int ld,
                                                      uint64_t* __restrict__ rng, float* __restrict__ actions, int n){
    extern __shared__ __align__(16) float smem[];
    float* img = smem;
    constexpr int ROWS = IN > MLP_HD ? IN : MLP_HD;
    float* scr = smem + MlpImg<IN, OUT>::SIZE + threadIdx.x;
    stage_mlp_image<IN, OUT>(img, blob, has_std != 0, has_log_std != 0);
    __syncthreads();
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    if(e >= n) return;
    for(int i = 0; i < IN; i++) scr[i * BLOCK] = obs[(size_t)e * ld + i];
    float o[OUT];
    mlp_forward<IN, OUT>(img, scr, BLOCK, o);
    if(head == B200L2F_HEAD_SQUASH_EVAL){
        for(int i = 0; i < 4; i++) actions[(size_t)e * 4 + i] = tanhf(o[i]);
    }
    else if(head == B200L2F_HEAD_PPO_GAUSSIAN){
        uint64_t s = rng[e];
        for(int i = 0; i < 4; i++) actions[(size_t)e * 4 + i] = rng_normal(s, o[i], expf(img[MlpImg<IN, OUT>::LOG_STD + i]));
        rng[e] = s;
    }
    else{
        for(int i = 0; i < 4; i++) actions[(size_t)e * 4 + i] = o[i];
    }
    (void)ROWS;
}

// ---------------------------------------------------------------------------------------------------------------
// closed-loop rollout with a deterministic MLP actor (head identity or squash-eval)
// ---------------------------------------------------------------------------------------------------------------
template <class Spec, int OUT>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_rollout_mlp(const __grid_constant__ RolloutArgs a, int has_std){
    constexpr int IN = Spec::OBS_DIM;
    constexpr int ROWS = IN > MLP_HD ? IN : MLP_HD;
    extern __shared__ __align__(16) float smem[];
    float* img = smem;
    float* sm_dyn = smem + MlpImg<IN, OUT>::SIZE;
    float* scr = sm_dyn + P_DYN_DIM * BLOCK + threadIdx.x;
    stage_mlp_image<IN, OUT>(img, a.blob, has_std != 0, false);
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    const bool active = e < a.n;
    const size_t n = (size_t)a.n;
    const size_t env = active ? (size_t)e : 0;
    ParamsStaged p = stage_dynamics(sm_dyn, a.params, n, env);
    __syncthreads();
    EnvState<Spec> st;
    load_state(st, a.state + env, n);
    DynInvariants d;
    dyn_invariants(d, p, st);
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = a.rng[env];
    float ret = 0.0f; int eplen = 0; bool done = false;
    for(int t = 0; t < a.T; t++){
        if(a.out_states && active && (t % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(t / a.state_stride) * n + env) * Spec::STATE_DIM);
        observe_to_scratch(st, p, rng, hist_ptr, n, scr, BLOCK);
        if(a.out_obs && active){
            float* row = a.out_obs + ((size_t)t * n + env) * IN;
            for(int i = 0; i < IN; i++) row[i] = scr[i * BLOCK];
        }
        float o[OUT], act[4];
        mlp_forward<IN, OUT>(img, scr, BLOCK, o);
#pragma unroll
        for(int i = 0; i < 4; i++) act[i] = OUT == 8 ? tanhf(o[i]) : o[i];
        if(a.out_actions && active) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env) * 4) = make_float4(act[0], act[1], act[2], act[3]);
        RewardInputs ri;
        reward_inputs(ri, st);
        if(Spec::H == 1 || active) env_step<Spec, true, ParamsStaged, true>(st, p, d, act, rng, hist_ptr, n);
        const bool term = env_terminated(p, st.x);
        const float r = env_reward(p, ri, act, st.x, term, d.dt);
        if(a.out_rewards && active) a.out_rewards[(size_t)t * n + env] = r;
        if(a.out_term && active) a.out_term[(size_t)t * n + env] = term ? 1 : 0;
        if(!done){ ret += r; eplen += 1; done = term; }
    }
    if(!active) return;
    if(a.out_states && (a.T % a.state_stride) == 0)
        write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(a.T / a.state_stride) * n + env) * Spec::STATE_DIM);
    store_state(st, a.state + env, n);
    a.rng[env] = rng;
    if(a.out_returns) a.out_returns[env] = ret;
    if(a.out_eplen) a.out_eplen[env] = eplen;
    if(a.out_done) a.out_done[env] = done ? 1 : 0;
    (void)ROWS;
}

// ---------------------------------------------------------------------------------------------------------------
// PPO collection (rl_tools::collect).  Dataset [(T+1)*n, D = OBS+15], row = step*n + env:
//   obs[OBS] | actions_mean[4] | actions[4] | log_prob | reward | terminated | truncated | value | advantage | target_value
// The last three columns belong to the learner and are not touched.  Rows of the 32 environments of a warp are consecutive in
// memory, so the warp stages its 32 x W floats in shared memory and writes them as contiguous 128-byte lines.
// ---------------------------------------------------------------------------------------------------------------
// floats of one warp's shared-memory slab in k_collect: [64 + IN][32] scratch columns (+ a separate [32][IN + 12] write-back window when the
// row does not fit in the 64 hidden-activation rows)
template <int IN> struct CollectSlab { static constexpr int FLOATS = (MLP_HD + IN) * 32 + (IN + 12 > MLP_HD ? 32 * (IN + 12) : 0); };
struct CollectArgs {
    float* params;            // [145][n]   (rewritten on reset)
    const float* env_row;     // [145] nominal / DR-range row the reset sampler starts from (device copy)
    float row[PARAMS_DIM];    // the same row by value: rides in the launch's constant bank (tensor-core kernel)
    float* state;             // [STATE_DIM][n] slot 0
    uint64_t* rng;
    const float* blob; int has_std;
    int* episode_step; float* episode_return; uint8_t* truncated;
    float* dataset;           // [(T+1)*n][D]
    int n, T, step_limit;
    int* error_flag;
    int bulk_rows;            // tensor-core kernel: dataset 16-byte aligned and n % 4 == 0 -> a warp's 32 rows leave as one bulk copy
    int n_chunks, chunk_steps; // tensor-core kernel: time-chunked scheduler (0 / 1 = one item per tile); the launcher enables it only for translation units built with ld.cg global loads
};
template <class Spec, bool DR>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_collect(const __grid_constant__ CollectArgs a){
    constexpr int IN = Spec::OBS_DIM, OUT = 4;
    constexpr int D = IN + 15, W = IN + 12;        // W: columns written per step
    constexpr int OBS0 = MLP_HD;                   // scratch rows [0, 64): hidden activations; [64, 64 + IN): the observation (kept for the write-back)
    constexpr int ROWS = MLP_HD + IN;              // per-warp slab [ROWS][32]; its first 32 * W floats double as the [32][W] write-back window
    constexpr bool WINDOW_APART = W > MLP_HD;      // ... unless the row is wider than the 64 hidden-activation rows (DEFAULT spec, OBS 82): own window
    constexpr int SLAB = CollectSlab<IN>::FLOATS;
    extern __shared__ __align__(16) float smem[];
    float* img = smem;
    float* sm_dyn = smem + MlpImg<IN, OUT>::SIZE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slab = sm_dyn + P_DYN_DIM * BLOCK + (size_t)warp * SLAB;        // private to this warp: only __syncwarp is needed
    float* scr = slab + lane;                                            // this thread's column, row stride 32
    float* win = WINDOW_APART ? slab + ROWS * 32 : slab;                 // [32][W] write-back window
    stage_mlp_image<IN, OUT>(img, a.blob, a.has_std != 0, true);
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    const bool active = e < a.n;
    const size_t n = (size_t)a.n;
    const size_t env = active ? (size_t)e : 0;
    ParamsStagedT<false> p = stage_dynamics<false>(sm_dyn, a.params, n, env);
    __syncthreads();
    EnvState<Spec> st;
    load_state(st, a.state + env, n);
    DynInvariants d;
    dyn_invariants(d, p, st);
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = a.rng[env];
    int ep_step = a.episode_step[env]; float ep_ret = a.episode_return[env]; bool truncated = a.truncated[env] != 0;
    const int warp_env0 = blockIdx.x * BLOCK + warp * 32;
    const int rows_valid = min(32, a.n - warp_env0);

    for(int t = 0; t <= a.T; t++){
        const bool last = t == a.T;                       // final observation only (operations_generic.h:122-129)
        if(!last && truncated && active){                 // prologue (operations_generic_per_env.h:17-25): re-sample parameters and state
            truncated = false; ep_step = 0; ep_ret = 0.0f;
            ParamsOverlay o;                              // sampled in registers: no dependent HBM round trips on the reset path
            o.init(a.env_row);
            if(!sample_parameters<DR, Spec::RNG_OOL>(o, rng)) atomicExch(a.error_flag, 1);
            o.flush(ParamsRW{a.params + env, n});
            p = stage_dynamics<false>(sm_dyn, a.params, n, env);   // re-stage this thread's column only
            sample_state<Spec, ParamsOverlay, true>(st, o, rng, hist_ptr, n);
            dyn_invariants(d, p, st);
        }
        observe_to_scratch(st, p, rng, hist_ptr, n, scr + OBS0 * 32, 32);
        float vals[12];
        if(!last){
            float mean[OUT], act[4];
            mlp_forward<IN, OUT, OBS0>(img, scr, 32, mean);
            float lp = 0.0f;
#pragma unroll
            for(int i = 0; i < 4; i++){                   // epilogue (operations_generic_per_env.h:43-58)
                const float ls = img[MlpImg<IN, OUT>::LOG_STD + i];
                act[i] = rng_normal_t<Spec::RNG_OOL>(rng, mean[i], expf(ls));
                lp += normal_log_prob(mean[i], ls, act[i]);
            }
            RewardInputs ri;
            reward_inputs(ri, st);
            if(Spec::H == 1 || active) env_step<Spec, true, ParamsStagedT<false>, true>(st, p, d, act, rng, hist_ptr, n);
            const bool term = env_terminated(p, st.x);
            const float r = env_reward(p, ri, act, st.x, term, d.dt);
            ep_ret += r; ep_step += 1;
            truncated = term || (a.step_limit > 0 && ep_step >= a.step_limit);
#pragma unroll
            for(int i = 0; i < 4; i++){ vals[i] = mean[i]; vals[4 + i] = act[i]; }
            vals[8] = lp; vals[9] = r; vals[10] = term ? 1.0f : 0.0f; vals[11] = truncated ? 1.0f : 0.0f;
        }
        // ---- coalesced write-back: every lane lays out ITS row in the [32][W] window (the hidden-activation rows are free now), then the warp
        // ---- streams the window to the 32 consecutive dataset rows as contiguous runs of W floats
        __syncwarp();
        for(int i = 0; i < IN; i++) win[lane * W + i] = scr[(OBS0 + i) * 32];
        if(!last){
#pragma unroll
            for(int i = 0; i < 12; i++) win[lane * W + IN + i] = vals[i];
        }
        __syncwarp();
        {
            float* gbase = a.dataset + ((size_t)t * n + warp_env0) * D;
            auto stream_rows = [&](auto ncols_c){          // compile-time divisor -> multiply-shift
                constexpr int NC = decltype(ncols_c)::value;
#pragma unroll 2
                for(int it = 0; it < NC; it++){
                    const int idx = lane + 32 * it;
                    const int r = idx / NC, c = idx - r * NC;
                    if(r < rows_valid) gbase[r * D + c] = win[r * W + c];
                }
            };
            if(!last) stream_rows(std::integral_constant<int, W>{});
            else stream_rows(std::integral_constant<int, IN>{});
        }
        __syncwarp();
    }
    if(!active) return;
    store_state(st, a.state + env, n);
    a.rng[env] = rng;
    a.episode_step[env] = ep_step; a.episode_return[env] = ep_ret; a.truncated[env] = truncated ? 1 : 0;
}

}  // namespace b200l2f
