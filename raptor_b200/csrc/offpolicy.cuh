// raptor_b200/csrc/offpolicy.cuh -- off-policy runner steps on device: the SAC-teacher data collection of the foundation-policy pre-training.
//
// Replaces rl::components::off_policy_runner `step` = prologue + interlude + epilogue
// (INC/rl/components/off_policy_runner/operations_generic.h:215-238; per environment: operations_generic_per_env.h:8-58 prologue_per_env,
// :60-110 epilogue_per_env; the reference's own CUDA kernels launch exactly these per-environment bodies, operations_cuda.h:62-106) and the
// replay buffer `add` (INC/rl/components/replay_buffer/operations_generic.h:54-79) for T runner steps in ONE launch:
//
//   per environment and step:  [reset if truncated: parameters (SAMPLE_PARAMETERS) + state re-sampled, ring bookkeeping]  -> observe
//                              -> SAC actor MLP IN-64-64-8 -> sample_and_squash in Mode<Rollout>: a = tanh(mean + N(0,1) exp(clamp(log_std)))
//                              -> step -> reward -> observe(next_state) -> terminated -> truncated = terminated | episode_step == limit
//                              -> replay row  obs[IN] | action[4] | reward | next_obs[IN] | terminated | truncated   at the ring position
//
// Data layout: every environment owns the reference's ReplayBuffer::data matrix, a ring [capacity][D = 2 IN + 7] (replay_buffer.h:37-58, views
// operations_generic.h:12-22; symmetric observations), stored back to back: replay[n][capacity][D]; episode_start[n][capacity]; position, full,
// current_episode_start[n].  Rows of neighbouring environments are capacity*D*4 bytes apart, so a warp transposes its 32 rows through a private
// shared-memory window and writes each row as one contiguous run: two phases of <= 32 floats (obs | action | reward, then next_obs | flags), one
// store instruction per row and phase.  RNG draw order per step = the reference's: reset samplers, observation noise, 4 exploration normals,
// action noise (step), next-observation noise.
#pragma once
#include "mlp.cuh"

namespace b200l2f {

struct OffPolicyArgs {
    CollectArgs c;            // params / env_row / row / state / rng / blob / episode_step / episode_return / truncated / n / T / step_limit / error_flag
    float* replay;            // [n][capacity][D]
    int* episode_start;       // [n][capacity]
    int* position; uint8_t* full; int* current_episode_start;   // [n]
    int capacity;
    int sample_parameters;    // OffPolicyRunner PARAMETERS::SAMPLE_PARAMETERS
};

// sample_and_squash evaluate_per_sample in Mode<Rollout> (INC/nn/layers/sample_and_squash/operations_generic.h:148-194), bounds layer.h:31-32
// FAST: default-math arithmetic (MUFU Box-Muller, ex2 / rcp exponential and tanh; the integer stream stays bit-exact)
template <bool OOL, bool FAST = false>
__device__ __forceinline__ void squash_sample(const float* __restrict__ o, uint64_t& rng, float* __restrict__ act){
#pragma unroll
    for(int i = 0; i < 4; i++){
        const float log_std = fminf(fmaxf(o[4 + i], -20.0f), 2.0f);
        const float noise = rng_normal_t<OOL, FAST>(rng, 0.0f, 1.0f);
        if constexpr(FAST) act[i] = tanh_of_scaled((2.0f * LOG2E) * (o[i] + noise * ex2_approx(log_std * LOG2E)));
        else act[i] = tanhf(o[i] + noise * expf(log_std));
    }
}

// one phase of the row write-back: every lane has put its NV values at win[lane * 33 + i]; row r of the warp goes to dst_r = base of lane r
template <int NV>
__device__ __forceinline__ void stream_phase(const float* __restrict__ win, float* my_dst, int lane, int rows_valid){
    static_assert(NV <= 32, "one store instruction per row");
    __syncwarp();
    const unsigned long long mine = reinterpret_cast<unsigned long long>(my_dst);
#pragma unroll 4
    for(int r = 0; r < 32; r++){
        float* dst = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, r));
        if(r < rows_valid && lane < NV) dst[lane] = win[r * 33 + lane];
    }
    __syncwarp();
}

// ---- shared per-step tail: ring bookkeeping after the row is written (replay_buffer/operations_generic.h:70-77)
struct Ring { int position, current_start; bool full; };
__device__ __forceinline__ void ring_advance(Ring& rg, int capacity, bool truncated){
    rg.position = rg.position + 1 == capacity ? 0 : rg.position + 1;
    if(truncated) rg.current_start = rg.position;
    if(rg.position == 0 && !rg.full) rg.full = true;
}

// ---------------------------------------------------------------------------------------------------------------
// gather_batch for SEQUENCE_LENGTH = 1 (the MLP SAC configuration): INC/rl/components/off_policy_runner/operations_generic.h:240-420 with the
// environment drawn per sample (:423-434; one RNG stream per batch sample like operations_cuda.h:36-60).  One warp per sample: lane 0 draws
// (environment, ring offset) -- uniform_int = next(state) % range, INC/random/operations_generic.h:43-50 -- the warp copies the 236-byte row.
// ---------------------------------------------------------------------------------------------------------------
struct GatherArgs {
    const float* replay; const int* position; const uint8_t* full;
    int obs_dim, capacity, max_episode_length, env_begin, env_count, batch;
    uint64_t* rng;
    float* observations_actions; float* rewards; uint8_t* terminated;
    uint8_t* reset; uint8_t* next_reset; uint8_t* final_step_mask; uint8_t* next_final_step_mask;
    int* env_index; int* sample_index;
    int* error_flag;
};
__global__ void __launch_bounds__(256) k_gather_batch(const GatherArgs a){
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(b >= a.batch) return;
    const int OBS = a.obs_dim, D = 2 * OBS + 7, W = OBS + 4;
    int env = 0, sample = -1;
    if(lane == 0){
        uint64_t s = a.rng[b];
        rng_next(s);
        env = a.env_begin + (int)(s % (uint64_t)a.env_count);
        const bool is_full = a.full[env] != 0;
        const int pos = a.position[env];
        const uint64_t eligible = is_full ? (uint64_t)a.capacity : (uint64_t)pos;
        if(eligible == 0) atomicExch(a.error_flag, 1);           // "Replay buffer requires at least one element" (:252)
        else{
            rng_next(s);
            const uint64_t offset = s % eligible;
            sample = is_full ? (int)(((uint64_t)pos + (uint64_t)a.max_episode_length + offset) % (uint64_t)a.capacity) : (int)offset;
        }
        a.rng[b] = s;
    }
    env = __shfl_sync(0xffffffffu, env, 0);
    sample = __shfl_sync(0xffffffffu, sample, 0);
    if(sample < 0) return;
    const float* row = a.replay + ((size_t)env * a.capacity + sample) * D;
    float* o0 = a.observations_actions + (size_t)b * W;
    float* o1 = a.observations_actions + ((size_t)a.batch + b) * W;
    for(int i = lane; i < W; i += 32){
        o0[i] = row[i];                                          // obs | action
        o1[i] = i < OBS ? row[OBS + 5 + i] : 0.0f;               // next_obs | next action = 0 (:408-410)
    }
    if(lane == 0){
        a.rewards[b] = row[OBS + 4];
        a.terminated[b] = row[2 * OBS + 5] != 0.0f ? 1 : 0;
        if(a.reset) a.reset[b] = 1;
        if(a.next_reset){ a.next_reset[b] = 1; a.next_reset[a.batch + b] = 1; }
        if(a.final_step_mask) a.final_step_mask[b] = 1;
        if(a.next_final_step_mask){ a.next_final_step_mask[b] = 0; a.next_final_step_mask[a.batch + b] = 1; }
        if(a.env_index) a.env_index[b] = env;
        if(a.sample_index) a.sample_index[b] = sample;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// gather_batch for any SEQUENCE_LENGTH (recurrent SAC; the batch parameters of off_policy_runner.h:78-85 as run-time values): gather_batch_step
// (operations_generic.h:240-420) per batch sample.  One warp per sample: lane 0 walks the padded sequence -- ring indices, the random sequence lengths
// (uniform_real in double = state * 2^-64, uniform_int = next(state) % range), the reset / final-step masks -- and broadcasts, per step, which ring
// row to copy and whether its next-observation opens the following (padding) step; the warp copies the rows.  rewards / terminated of padding steps,
// which the reference leaves unwritten, are zero.
// ---------------------------------------------------------------------------------------------------------------
struct GatherSeqArgs {
    GatherArgs g;                 // rings, rng, outputs (the masks are mandatory here), environment range, batch
    const int* episode_start;     // [n][capacity]
    int L, include_first, always_initial, random_len, enable_nominal;
    float nominal_probability;
};
__device__ __forceinline__ double rng_unit_double(uint64_t& s){ rng_next(s); return __ull2double_rn(s) * 5.42101086242752217e-20; }   // state / (double)MAX_INDEX, exact scaling by 2^-64
__global__ void __launch_bounds__(256) k_gather_batch_sequential(const GatherSeqArgs q){
    const GatherArgs& a = q.g;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if(b >= a.batch) return;
    const int OBS = a.obs_dim, D = 2 * OBS + 7, W = OBS + 4, L = q.L, P = L + 1, B = a.batch;
    // lane 0's walk state
    uint64_t s = 0;
    int env = 0, pos = 0, sample_index = 0, first_sample = -1;
    unsigned eligible = 0, cur_len = (unsigned)L, cur_step = 0;
    bool is_full = false, prev_trunc = false, prev_pad = true, ok = true;
    auto draw_length = [&]() -> unsigned {   // :265-274, :361-370
        if(q.enable_nominal && rng_unit_double(s) < (double)q.nominal_probability) return (unsigned)L;
        if(L > 1){ rng_next(s); return 1u + (unsigned)(s % (uint64_t)(L - 1)); }
        return (unsigned)L;
    };
    if(lane == 0){
        s = a.rng[b];
        rng_next(s);
        env = a.env_begin + (int)(s % (uint64_t)a.env_count);
        is_full = a.full[env] != 0; pos = a.position[env];
        const int el = is_full ? (q.always_initial ? a.capacity - a.max_episode_length : a.capacity) : pos;
        if(el < 1){ atomicExch(a.error_flag, 1); ok = false; }      // "Replay buffer requires at least one element" (:252-256)
        eligible = (unsigned)max(el, 1);
        if(q.random_len) cur_len = draw_length();
        for(int t = 0; t < L; t++){ a.final_step_mask[(size_t)t * B + b] = 0; a.reset[(size_t)t * B + b] = 0; a.rewards[(size_t)t * B + b] = 0.0f; a.terminated[(size_t)t * B + b] = 0; }
        for(int t = 0; t < P; t++){ a.next_final_step_mask[(size_t)t * B + b] = 0; a.next_reset[(size_t)t * B + b] = 0; }
        a.reset[b] = 1; a.next_reset[b] = 1; a.next_reset[(size_t)(q.include_first ? 0 : 1) * B + b] = 1;   // :286-288 (the `next_reset` view starts at row 0 or 1)
    }
    ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
    env = __shfl_sync(0xffffffffu, env, 0);
    if(!ok) return;
    const float* ring = a.replay + (size_t)env * a.capacity * D;
    for(int seq = 0; seq < P; seq++){
        int row_i = -1; bool truncated = false;                      // row_i < 0: padding step, nothing to copy
        if(lane == 0){
            if(prev_trunc){                                          // :290-299
                prev_trunc = false; prev_pad = true;
                a.next_final_step_mask[(size_t)seq * B + b] = 1;
                if(seq < P - 1) a.reset[(size_t)seq * B + b] = 1;
            }
            else{
                if(prev_pad){                                        // a new sequence starts here (:300-321)
                    if(seq < P - 1) a.reset[(size_t)seq * B + b] = 1;
                    a.next_reset[(size_t)seq * B + b] = 1;
                    rng_next(s);
                    const unsigned offset = (unsigned)(s % (uint64_t)eligible);
                    sample_index = is_full ? (int)(((uint64_t)pos + (uint64_t)a.max_episode_length + offset) % (uint64_t)a.capacity) : (int)offset;
                    if(q.always_initial) sample_index = q.episode_start[(size_t)env * a.capacity + sample_index];
                    if(first_sample < 0) first_sample = sample_index;
                }
                row_i = sample_index;
                const float* row = ring + (size_t)row_i * D;
                if(seq < P - 1){ a.rewards[(size_t)seq * B + b] = row[OBS + 4]; a.terminated[(size_t)seq * B + b] = row[2 * OBS + 5] != 0.0f ? 1 : 0; }
                truncated = row[2 * OBS + 6] != 0.0f;                // :352-360
                int next = sample_index + 1;
                if(is_full) next = next % a.capacity;
                if(next == pos) truncated = true;
                if(seq == P - 2) truncated = true;
                if(q.random_len){                                    // :361-388
                    if(cur_step == cur_len - 1){
                        truncated = true;
                        cur_len = L > 1 ? draw_length() : (unsigned)L;
                    }
                    cur_step = truncated ? 0 : cur_step + 1;
                }
                if(truncated && seq < P - 1) a.final_step_mask[(size_t)seq * B + b] = 1;
                sample_index = next;
                prev_pad = false; prev_trunc = truncated;
            }
        }
        row_i = __shfl_sync(0xffffffffu, row_i, 0);
        truncated = __shfl_sync(0xffffffffu, (int)truncated, 0) != 0;
        if(row_i < 0) continue;
        const float* row = ring + (size_t)row_i * D;
        float* oa = a.observations_actions + ((size_t)seq * B + b) * W;
        const bool open_next = truncated && seq < P - 1;             // next_obs | action 0 into the following step (:393-417)
        for(int i = lane; i < W; i += 32){
            oa[i] = row[i];
            if(open_next) oa[(size_t)B * W + i] = i < OBS ? row[OBS + 5 + i] : 0.0f;
        }
    }
    if(lane == 0){
        a.rng[b] = s;
        if(a.env_index) a.env_index[b] = env;
        if(a.sample_index) a.sample_index[b] = first_sample;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// actor on fp32 CUDA cores (B200L2F_GEMM_FP32_CUDA_CORES, B200L2F_FLAG_ACCURATE_MATH): structure of k_collect (mlp.cuh)
// ---------------------------------------------------------------------------------------------------------------
template <class Spec, bool DR>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_off_policy(const __grid_constant__ OffPolicyArgs oa){
    const CollectArgs& a = oa.c;
    constexpr int IN = Spec::OBS_DIM, OUT = 8;
    constexpr int D = 2 * IN + 7;
    constexpr int OBS0 = MLP_HD;                   // scratch rows [0, 64): hidden activations / write-back window; [64, 64 + IN): the observation
    constexpr int ROWS = MLP_HD + IN;
    static_assert(33 * 32 <= MLP_HD * 32, "the [32][33] window must fit in the hidden-activation rows");
    extern __shared__ __align__(16) float smem[];
    float* img = smem;
    float* sm_dyn = smem + MlpImg<IN, OUT>::SIZE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slab = sm_dyn + P_DYN_DIM * BLOCK + (size_t)warp * ROWS * 32;
    float* scr = slab + lane;
    stage_mlp_image<IN, OUT>(img, a.blob, a.has_std != 0, false);
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
    Ring rg{oa.position[env], oa.current_episode_start[env], oa.full[env] != 0};
    float* ring = oa.replay + env * (size_t)oa.capacity * D;
    int* es = oa.episode_start + env * (size_t)oa.capacity;
    const int warp_env0 = (int)blockIdx.x * BLOCK + warp * 32;   // signed: a.n - warp_env0 is negative for a fully inactive warp
    const int rows_valid = min(32, a.n - warp_env0);

    for(int t = 0; t < a.T; t++){
        if(truncated && active){                          // prologue_per_env (operations_generic_per_env.h:20-57)
            truncated = false; ep_step = 0; ep_ret = 0.0f;
            if(oa.sample_parameters){
                ParamsOverlay o;
                o.init(a.env_row);
                if(!sample_parameters<DR, Spec::RNG_OOL>(o, rng)) atomicExch(a.error_flag, 1);
                o.flush(ParamsRW{a.params + env, n});
                p = stage_dynamics<false>(sm_dyn, a.params, n, env);
                sample_state<Spec, ParamsOverlay, true>(st, o, rng, hist_ptr, n);
            }
            else sample_state<Spec, ParamsStagedT<false>, true>(st, p, rng, hist_ptr, n);
            dyn_invariants(d, p, st);
            if(rg.full || rg.position > 0){
                const int previous = rg.position == 0 ? oa.capacity - 1 : rg.position - 1;
                ring[(size_t)previous * D + D - 1] = 1.0f;
                rg.current_start = rg.position;
            }
        }
        observe_to_scratch(st, p, rng, hist_ptr, n, scr + OBS0 * 32, 32);
        float o8[OUT], act[4];
        mlp_forward<IN, OUT, OBS0>(img, scr, 32, o8);
        squash_sample<Spec::RNG_OOL>(o8, rng, act);        // interlude: evaluate_step in Mode<Rollout>
        RewardInputs ri;                                  // epilogue_per_env (:60-110)
        reward_inputs(ri, st);
        if(Spec::H == 1 || active) env_step<Spec, true, ParamsStagedT<false>, true>(st, p, d, act, rng, hist_ptr, n);
        const bool term = env_terminated(p, st.x);
        const float r = env_reward(p, ri, act, st.x, term, d.dt);
        ep_ret += r; ep_step += 1;
        truncated = term || ep_step == a.step_limit;
        float* row = ring + (size_t)rg.position * D;
        // phase 1: obs | action | reward
        __syncwarp();
        for(int i = 0; i < IN; i++) slab[lane * 33 + i] = scr[(OBS0 + i) * 32];
#pragma unroll
        for(int i = 0; i < 4; i++) slab[lane * 33 + IN + i] = act[i];
        slab[lane * 33 + IN + 4] = r;
        stream_phase<IN + 5>(slab, row, lane, rows_valid);
        // phase 2: next_obs | terminated | truncated
        observe_to_scratch(st, p, rng, hist_ptr, n, scr + OBS0 * 32, 32);
        __syncwarp();
        for(int i = 0; i < IN; i++) slab[lane * 33 + i] = scr[(OBS0 + i) * 32];
        slab[lane * 33 + IN] = term ? 1.0f : 0.0f;
        slab[lane * 33 + IN + 1] = truncated ? 1.0f : 0.0f;
        stream_phase<IN + 2>(slab, row + IN + 5, lane, rows_valid);
        if(active) es[rg.position] = rg.current_start;
        ring_advance(rg, oa.capacity, truncated);
    }
    if(!active) return;
    store_state(st, a.state + env, n);
    a.rng[env] = rng;
    a.episode_step[env] = ep_step; a.episode_return[env] = ep_ret; a.truncated[env] = truncated ? 1 : 0;
    oa.position[env] = rg.position; oa.current_episode_start[env] = rg.current_start; oa.full[env] = rg.full ? 1 : 0;
}

}  // namespace b200l2f
