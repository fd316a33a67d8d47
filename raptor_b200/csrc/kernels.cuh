// raptor_b200/csrc/kernels.cuh -- the engine's kernels (sm_100a).  One environment per thread, 128-thread CTAs.
//
// Vector-API kernels (one reference call each) and the fused persistent rollout kernel that replaces the loop body of
// rl_tools::evaluate (rl_tools/rl/utils/evaluation/operations_generic.h:138-189): observe -> actor -> step -> reward -> terminated
// for T steps in ONE launch with the integrated state, the GRU hidden state and the RNG stream resident in registers, the
// per-environment dynamics parameters staged once in shared memory and the actor weights staged once in shared memory.
#pragma once
#include <cstdint>
#include "layout.h"
#include "rng.cuh"
#include "env.cuh"
#include "samplers.cuh"
#include "policy.cuh"

namespace b200l2f {

#ifndef B200L2F_BLOCK
#define B200L2F_BLOCK 128
#endif
#ifndef B200L2F_MIN_BLOCKS
#define B200L2F_MIN_BLOCKS 2
#endif
constexpr int BLOCK = B200L2F_BLOCK;               // environments (= threads) per CTA
constexpr int MIN_BLOCKS = B200L2F_MIN_BLOCKS;     // resident CTAs per SM the fused kernels are register-budgeted for

// stage the dynamics block of this thread's environment: sm[i * BLOCK + tid] = params[i][env]; time constants -> reciprocals
template <bool NC = true>
__device__ __forceinline__ ParamsStagedT<NC> stage_dynamics(float* __restrict__ sm_dyn, const float* params, size_t n, size_t env){
    float* sm = sm_dyn + threadIdx.x;
    const float* g = params + env;
#pragma unroll 4
    for(int i = 0; i < P_DYN_DIM; i++) sm[i * BLOCK] = NC ? __ldg(g + (size_t)i * n) : g[(size_t)i * n];
#pragma unroll
    for(int r = 0; r < 8; r++) sm[(P_TAU_RISE + r) * BLOCK] = 1.0f / sm[(P_TAU_RISE + r) * BLOCK];
    ParamsStagedT<NC> p;
    p.sm = sm; p.sm_stride = BLOCK; p.base = g; p.stride = n;
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// utility kernels
// ---------------------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c]; 32x32 tiles through shared memory, coalesced on both sides (AoS rows <-> SoA columns)
static __global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int rows, int cols){
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for(int j = threadIdx.y; j < 32; j += blockDim.y){
        const int r = r0 + j, c = c0 + threadIdx.x;
        if(r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for(int j = threadIdx.y; j < 32; j += blockDim.y){
        const int c = c0 + j, r = r0 + threadIdx.x;
        if(r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}
static __global__ void k_init_rng(uint64_t* __restrict__ rng, int n, uint64_t seed, uint64_t first_env, int warmup){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    uint64_t s = rng_seed_state(seed + first_env + (uint64_t)e);
    for(int i = 0; i < warmup; i++) rng_next(s);
    rng[e] = s;
}
static __global__ void k_fill_params(float* __restrict__ params, const float* __restrict__ env_row, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    for(int i = 0; i < PARAMS_DIM; i++) params[(size_t)i * n + e] = env_row[i];
}
template <bool DR>
__global__ void k_sample_params(float* __restrict__ params, const float* __restrict__ env_row, uint64_t* __restrict__ rng, int n, int* __restrict__ error_flag){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    uint64_t s = rng[e];
    ParamsRW p{params + e, (size_t)n};
    if(!sample_parameters<DR>(env_row, p, s)) atomicExch(error_flag, 1);
    rng[e] = s;
}
// MDP parameters that are normally identical for every environment (reward weights, termination switches, Langevin constants): when they
// are, the fused kernels read them from the launch's constant bank instead of per-environment loads inside the time loop.
__host__ __device__ constexpr bool mdp_uniform_index(int i){
    return (i >= P_RW_NONNEG && i <= P_RW_POS_INTEGRAL) || i == P_HOVER || i == P_TERM_ENABLED || i == P_TERM_LINVEL || i == P_TERM_ANGVEL ||
           (i >= P_LANGEVIN_GAMMA && i <= P_LANGEVIN_ALPHA) || (i >= P_NOISE_POS && i <= P_ACTION_NOISE);
}
// features of the parameter set that select kernel variants: bit0 = some observation/action noise std != 0,
// bit1 = some environment's "uniform" MDP parameter differs from environment 0's, bit2 = some vehicle is not "axial" (a rotor thrust
// direction other than body z, or an off-diagonal entry in J / J^-1): the fused kernels then keep the general rotor / inertia matrices
static __global__ void k_param_features(const float* __restrict__ params, int n, int* __restrict__ features){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    int f = 0;
    if(e < n){
        auto P = [&](int i){ return params[(size_t)i * n + e]; };
        for(int i = P_NOISE_POS; i <= P_ACTION_NOISE; i++) if(P(i) != 0.0f) f |= 1;
        for(int i = 0; i < PARAMS_DIM; i++) if(mdp_uniform_index(i) && P(i) != params[(size_t)i * n]) f |= 2;
        for(int r = 0; r < 4; r++) if(P(P_THRUST_DIR + 3 * r) != 0.0f || P(P_THRUST_DIR + 3 * r + 1) != 0.0f || P(P_THRUST_DIR + 3 * r + 2) != 1.0f) f |= 4;
        for(int i = 0; i < 9; i++) if(i % 4 != 0 && (P(P_J + i) != 0.0f || P(P_JINV + i) != 0.0f)) f |= 4;
    }
    f = __reduce_or_sync(0xffffffffu, f);
    if((threadIdx.x & 31) == 0 && f) atomicOr(features, f);
}

// ---------------------------------------------------------------------------------------------------------------
// status of the last fused call: non-finite environments (L2F/operations_generic/05_state_is_nan.h: any NaN in the state; Inf counted as well)
// and the aggregates of rl::utils::evaluation::Result (operations_generic.h:201-213).  Two deterministic passes: per-block partial sums in
// double, then one block adds the partials in block order.
// ---------------------------------------------------------------------------------------------------------------
struct StatusPartial { double ret, ret2, len, len2; unsigned long long terminated, nonfinite; };
static __global__ void __launch_bounds__(256) k_status_partials(const float* __restrict__ state, int sdim, int n, const float* __restrict__ returns, const int* __restrict__ eplen,
                                                                  const uint8_t* __restrict__ done, uint8_t* __restrict__ nonfinite_flags, StatusPartial* __restrict__ partials){
    double ret = 0, ret2 = 0, len = 0, len2 = 0; unsigned long long term = 0, bad = 0;
    for(int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x){
        bool finite = true;
        for(int i = 0; i < sdim; i++) finite = finite && isfinite(state[(size_t)i * n + e]);
        if(nonfinite_flags) nonfinite_flags[e] = finite ? 0 : 1;
        bad += finite ? 0 : 1;
        if(returns){ const double r = returns[e]; ret += r; ret2 += r * r; }
        if(eplen){ const double l = eplen[e]; len += l; len2 += l * l; }
        if(done) term += done[e] ? 1 : 0;
    }
    __shared__ StatusPartial sm[256];
    sm[threadIdx.x] = StatusPartial{ret, ret2, len, len2, term, bad};
    __syncthreads();
    for(int s = 128; s > 0; s >>= 1){
        if((int)threadIdx.x < s){
            StatusPartial& a = sm[threadIdx.x]; const StatusPartial& b = sm[threadIdx.x + s];
            a.ret += b.ret; a.ret2 += b.ret2; a.len += b.len; a.len2 += b.len2; a.terminated += b.terminated; a.nonfinite += b.nonfinite;
        }
        __syncthreads();
    }
    if(threadIdx.x == 0) partials[blockIdx.x] = sm[0];
}
static __global__ void k_status_finish(const StatusPartial* __restrict__ partials, int n_partials, StatusPartial* __restrict__ out){
    if(threadIdx.x != 0 || blockIdx.x != 0) return;
    StatusPartial t{0, 0, 0, 0, 0, 0};
    for(int i = 0; i < n_partials; i++){ const StatusPartial& b = partials[i]; t.ret += b.ret; t.ret2 += b.ret2; t.len += b.len; t.len2 += b.len2; t.terminated += b.terminated; t.nonfinite += b.nonfinite; }
    *out = t;
}

// features + parameter row of environment 0 -> page-locked host words, written by the device itself (zero-copy store over PCIe): a cudaMemcpy D2H would queue
// on the copy engine behind whatever download is in flight (measured: 0.25 ms behind a 12.6 MB state download, per pipelined rollout)
static __global__ void k_publish_features(const int* __restrict__ features, const float* __restrict__ params, int n, int* __restrict__ host_features, float* __restrict__ host_row0){
    const int i = threadIdx.x;
    if(i < PARAMS_DIM) host_row0[i] = params[(size_t)i * n];
    if(i == 0) *host_features = *features;
    __threadfence_system();
}

template <class Spec, bool SAMPLE>
__global__ void k_init_state(const float* __restrict__ params, float* __restrict__ state, uint64_t* __restrict__ rng, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    ParamsGlobal p{params + e, (size_t)n};
    EnvState<Spec> st;
    float* hist = state + (size_t)S_HIST * n + e;
    if constexpr(SAMPLE){
        uint64_t s = rng[e];
        sample_state(st, p, s, hist, (size_t)n);
        rng[e] = s;
    }
    else{
        initial_state(st, p, hist, (size_t)n);
    }
    store_state(st, state + e, (size_t)n);
}

template <class Spec>
__global__ void k_observe(const float* __restrict__ params, const float* __restrict__ state, uint64_t* __restrict__ rng, float* __restrict__ obs, int ld, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    ParamsGlobal p{params + e, (size_t)n};
    EnvState<Spec> st;
    load_state(st, state + e, (size_t)n);
    uint64_t s = rng[e];
    float o[18];
    observe18<Spec, true>(st, p, s, o);
    rng[e] = s;
    float* row = obs + (size_t)e * ld;
#pragma unroll
    for(int i = 0; i < 18; i++) row[i] = o[i];
    float tail[4 * Spec::H + 4];
    observe_tail(st, p, state + (size_t)S_HIST * n + e, (size_t)n, tail, Spec::H);
    for(int i = 0; i < Spec::OBS_DIM - 18; i++) row[18 + i] = tail[i];
}

template <class Spec>
__global__ void __launch_bounds__(BLOCK) k_step(const float* __restrict__ params, const float* __restrict__ state, const float* __restrict__ actions,
                                                  float* __restrict__ next, uint64_t* __restrict__ rng, float* __restrict__ dts, int n){
    extern __shared__ float sm_dyn[];
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    if(e >= n) return;
    ParamsStaged p = stage_dynamics(sm_dyn, params, (size_t)n, (size_t)e);
    EnvState<Spec> st;
    load_state(st, state + e, (size_t)n);
    DynInvariants d;
    dyn_invariants(d, p, st);
    float a[4];
    const float4 av = *reinterpret_cast<const float4*>(actions + (size_t)e * 4);
    a[0] = av.x; a[1] = av.y; a[2] = av.z; a[3] = av.w;
    uint64_t s = rng[e];
    float* hist_next = next + (size_t)S_HIST * n + e;
    if constexpr(Spec::H > 1){
        if(next != state){
            const float* hist_prev = state + (size_t)S_HIST * n + e;
            for(int i = 0; i < 4 * Spec::H; i++) hist_next[(size_t)i * n] = hist_prev[(size_t)i * n];
        }
    }
    env_step<Spec, true>(st, p, d, a, s, hist_next, (size_t)n);
    rng[e] = s;
    store_state(st, next + e, (size_t)n);
    if(dts) dts[e] = d.dt;
}

template <class Spec>
__global__ void k_reward(const float* __restrict__ params, const float* __restrict__ state, const float* __restrict__ actions, const float* __restrict__ next,
                         float* __restrict__ rewards, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    ParamsGlobal p{params + e, (size_t)n};
    EnvState<Spec> st, nx;
    load_state(st, state + e, (size_t)n);
    load_state(nx, next + e, (size_t)n);
    RewardInputs ri;
    reward_inputs(ri, st);
    float a[4];
#pragma unroll
    for(int i = 0; i < 4; i++) a[i] = actions[(size_t)e * 4 + i];
    rewards[e] = env_reward(p, ri, a, nx.x, env_terminated(p, nx.x), p[P_DT]);
}
template <class Spec>
__global__ void k_terminated(const float* __restrict__ params, const float* __restrict__ state, uint8_t* __restrict__ flags, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    ParamsGlobal p{params + e, (size_t)n};
    EnvState<Spec> st;
    load_state(st, state + e, (size_t)n);
    flags[e] = env_terminated(p, st.x) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// actor kernels (Raptor GRU).  hidden is SoA [HD][n].
// ---------------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void k_policy_reset(float* __restrict__ hidden, int* __restrict__ gru_step, const float* __restrict__ h0, const uint8_t* __restrict__ mask, int n){
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if(e >= n) return;
    if(mask && !mask[e]) return;
    for(int j = 0; j < HD; j++) hidden[(size_t)j * n + e] = h0[j];
    gru_step[e] = 0;
}
template <int IN, int HD, int OUT, bool FAST>
__global__ void __launch_bounds__(BLOCK) k_raptor_step(const float* __restrict__ blob, const float* __restrict__ obs, int ld, float* __restrict__ hidden, int* __restrict__ gru_step,
                                                         int seq_len, int no_auto_reset, float* __restrict__ actions, int n){
    extern __shared__ __align__(16) float sm_img[];
    stage_raptor<IN, HD, OUT>(sm_img, blob);
    __syncthreads();
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    if(e >= n) return;
    float o[IN], h[HD], a[OUT];
#pragma unroll
    for(int i = 0; i < IN; i++) o[i] = obs[(size_t)e * ld + i];
#pragma unroll
    for(int j = 0; j < HD; j++) h[j] = hidden[(size_t)j * n + e];
    int gs = gru_step[e];
    raptor_forward<IN, HD, OUT, FAST>(WeightsShared{sm_img}, o, h, gs, seq_len, no_auto_reset != 0, a);
#pragma unroll
    for(int j = 0; j < HD; j++) hidden[(size_t)j * n + e] = h[j];
    gru_step[e] = gs;
#pragma unroll
    for(int j = 0; j < OUT; j++) actions[(size_t)e * OUT + j] = a[j];
}

// ---------------------------------------------------------------------------------------------------------------
// THE hot path: fused persistent rollout with the Raptor actor
// ---------------------------------------------------------------------------------------------------------------
struct RolloutArgs {
    const float* params;   // [145][n]
    float* state;          // [STATE_DIM][n], advanced in place
    uint64_t* rng;         // [n]
    float* hidden;         // [HD][n]
    int* gru_step;         // [n]
    const float* blob;     // actor weights (row-major blob)
    int n, T, no_auto_reset, seq_len;
    // optional outputs
    float* out_states; int state_stride;  // [T/stride + 1][n][STATE_DIM]
    float* out_obs;        // [T][n][IN]
    float* out_actions;    // [T][n][4]
    float* out_rewards;    // [T][n]
    uint8_t* out_term;     // [T][n]
    float* out_returns;    // [n]
    int* out_eplen;        // [n]
    uint8_t* out_done;     // [n] the episode terminated within the rollout (rl/utils/evaluation: num_terminated)
    float row0[PARAMS_DIM];  // parameter row of environment 0 (source of the uniform MDP constants)
    // time-chunked persistent scheduler (tensor-core TS kernel): sched[0] = work counter, sched[1 + tile] = chunks published for the tile
    int* sched; int n_chunks, chunk_steps;
    float* acc_ret; int* acc_len;   // per-environment return / (episode length << 1 | done) carried between the chunks of one rollout
};

template <class Spec>
__device__ __forceinline__ void write_state_row(const EnvState<Spec>& st, const float* __restrict__ hist_ptr, size_t n, float* __restrict__ row){
#pragma unroll
    for(int i = 0; i < 13; i++) row[i] = st.x[i];
#pragma unroll
    for(int i = 0; i < 4; i++) row[S_LAST_ACTION + i] = st.last_action[i];
#pragma unroll
    for(int i = 0; i < 3; i++) row[S_ANGVEL_HIST + i] = st.x[X_OMEGA + i];
#pragma unroll
    for(int i = 0; i < 3; i++){ row[S_FORCE + i] = st.force[i]; row[S_TORQUE + i] = st.torque[i]; }
#pragma unroll
    for(int i = 0; i < 4; i++) row[S_RPM + i] = st.x[X_RPM + i];
    row[S_CURRENT_STEP] = (float)st.current_step;
    if constexpr(Spec::H == 1){
#pragma unroll
        for(int i = 0; i < 4; i++) row[S_HIST + i] = st.hist[i];
    }
    else{
        for(int i = 0; i < 4 * Spec::H; i++) row[S_HIST + i] = hist_ptr[(size_t)i * n];
    }
    row[s_traj_type(Spec::H)] = (float)st.traj_type;
    if constexpr(Spec::LANGEVIN){
#pragma unroll
        for(int i = 0; i < 12; i++) row[s_langevin(Spec::H) + i] = st.lang[i];
    }
    else{
#pragma unroll
        for(int i = 0; i < 12; i++) row[s_langevin(Spec::H) + i] = 0.0f;
    }
}

// CONSTW: actor weights in the launch's constant bank instead of shared memory.  ROLLED: compact-code variant (rolled actor k-loops with the
// loop-indexed vectors and the hidden state in a shared-memory scratch column, RK4 stages as one loop body).
template <class Spec, int IN, int HD, int OUT, bool NOISE, bool FAST, bool CONSTW, bool ROLLED>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_rollout_raptor(const __grid_constant__ RolloutArgs a, const __grid_constant__ WeightBlock<CONSTW ? RaptorImage<IN, HD, OUT>::SIZE : 1> wk){
    static_assert(!(CONSTW && ROLLED), "the rolled actor indexes the weight image with a loop variable: it needs the shared-memory image");
    static_assert(IN == 22 && OUT == 4, "the Raptor actor consumes the first 22 observation columns and emits 4 motor commands");
    extern __shared__ __align__(16) float smem[];
    constexpr int IMG = RaptorImage<IN, HD, OUT>::SIZE;
    float* sm_img = smem;                                         // IMG floats (unused when the weights ride in the constant bank)
    float* sm_dyn = smem + (CONSTW ? 0 : IMG);                    // P_DYN_DIM * BLOCK floats
    using SCR = RaptorScratch<IN, HD>;
    float* scr = sm_dyn + P_DYN_DIM * BLOCK + threadIdx.x;       // ROLLED: SCR::ROWS * BLOCK floats, this thread's column
    if constexpr(!CONSTW) stage_raptor<IN, HD, OUT>(sm_img, a.blob);
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    const bool active = e < a.n;
    const size_t n = (size_t)a.n;
    const size_t env = active ? (size_t)e : 0;                    // inactive lanes shadow environment 0 and never store
    ParamsStaged p = stage_dynamics(sm_dyn, a.params, n, env);
    __syncthreads();
    EnvState<Spec> st;
    load_state(st, a.state + env, n);
    DynInvariants d;
    dyn_invariants(d, p, st);
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = a.rng[env];
    float h[ROLLED ? 1 : HD];
    if constexpr(ROLLED){
#pragma unroll
        for(int j = 0; j < HD; j++) scr[(SCR::H + j) * BLOCK] = a.hidden[(size_t)j * n + env];
    }
    else{
#pragma unroll
        for(int j = 0; j < HD; j++) h[j] = a.hidden[(size_t)j * n + env];
    }
    int gs = a.gru_step[env];
    float ret = 0.0f; int eplen = 0; bool done = false;
    const bool no_auto_reset = a.no_auto_reset != 0;

    for(int t = 0; t < a.T; t++){
        if(a.out_states && active && (t % a.state_stride) == 0)
            write_state_row(st, hist_ptr, n, a.out_states + ((size_t)(t / a.state_stride) * n + env) * Spec::STATE_DIM);
        float obs[IN];
        observe18<Spec, NOISE>(st, p, rng, obs);
        if constexpr(Spec::H == 1){
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = st.hist[i];
        }
        else{   // most recent ring entry (40_observe.h:281); before the first step it holds the normalised initial rpm
            const int cur = st.current_step == 0 ? Spec::H - 1 : st.current_step - 1;
#pragma unroll
            for(int i = 0; i < 4; i++) obs[18 + i] = hist_ptr[(size_t)(4 * cur + i) * n];
        }
        if(a.out_obs && active){
            float* row = a.out_obs + ((size_t)t * n + env) * IN;
#pragma unroll
            for(int i = 0; i < IN; i++) row[i] = obs[i];
        }
        float act[OUT];
        if constexpr(ROLLED){
#pragma unroll
            for(int i = 0; i < IN; i++) scr[(SCR::OBS + i) * BLOCK] = obs[i];
            raptor_forward_rolled<IN, HD, OUT, FAST>(sm_img, scr, BLOCK, gs, a.seq_len, no_auto_reset, act);
        }
        else if constexpr(CONSTW) raptor_forward<IN, HD, OUT, FAST>(WeightsParam<IMG>{wk}, obs, h, gs, a.seq_len, no_auto_reset, act);
        else raptor_forward<IN, HD, OUT, FAST>(WeightsShared{sm_img}, obs, h, gs, a.seq_len, no_auto_reset, act);
        if(a.out_actions && active) *reinterpret_cast<float4*>(a.out_actions + ((size_t)t * n + env) * 4) = make_float4(act[0], act[1], act[2], act[3]);
        RewardInputs ri;
        reward_inputs(ri, st);
        if(Spec::H == 1 || active) env_step<Spec, NOISE, ParamsStaged, ROLLED>(st, p, d, act, rng, hist_ptr, n);   // H > 1 writes the ring in HBM: shadow lanes must not
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
#pragma unroll
    for(int j = 0; j < HD; j++) a.hidden[(size_t)j * n + env] = ROLLED ? scr[(SCR::H + j) * BLOCK] : h[ROLLED ? 0 : j];
    a.gru_step[env] = gs;
    if(a.out_returns) a.out_returns[env] = ret;
    if(a.out_eplen) a.out_eplen[env] = eplen;
    if(a.out_done) a.out_done[env] = done ? 1 : 0;
}

}  // namespace b200l2f
