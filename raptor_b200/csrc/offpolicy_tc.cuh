// raptor_b200/csrc/offpolicy_tc.cuh -- k_off_policy_ts: the off-policy runner steps of offpolicy.cuh with the SAC actor's three GEMMs on tcgen05
// (3xTF32, A operand and accumulators in TMEM, weight image by TMA: mlp_tc.cuh).  Structure of k_collect_ts: persistent tile loop, an
// environment's reset is divergent CUDA-core work inside the step loop, the GEMMs are CTA-collective and sit outside every lane-dependent branch.
#pragma once
#include "mlp_tc.cuh"
#include "offpolicy.cuh"

namespace b200l2f {

template <int IN>
struct OffPolicyTsSmem {
    using SM = MlpTsSmem<IN, 8>;
    static constexpr int WINDOW = SM::SLAB;                      // per-warp [32][33] write-back windows
    static constexpr int TOTAL = WINDOW + 4 * 32 * 33 * 4;
};

template <class Spec, bool DR, bool FOLLOW, bool AXIAL>
__global__ void __launch_bounds__(BLOCK, 2) k_off_policy_ts(const __grid_constant__ OffPolicyArgs oa, const float* __restrict__ tc_image, int* __restrict__ sched){
    const CollectArgs& a = oa.c;
    constexpr int IN = Spec::OBS_DIM, OUT = 8;
    constexpr int D = 2 * IN + 7;
    using SM = MlpTsSmem<IN, OUT>;
    extern __shared__ __align__(1024) unsigned char smraw[];
    float* sm_dyn = reinterpret_cast<float*>(smraw + SM::DYN);
    TsCtx c = mlp_ts_prologue<IN, OUT>(smraw, tc_image);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* win = reinterpret_cast<float*>(smraw + OffPolicyTsSmem<IN>::WINDOW) + (size_t)warp * 32 * 33;   // private to this warp
    const size_t n = (size_t)a.n;
    __shared__ int s_item;
    const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
    for(;;){
    if(tid == 0) s_item = atomicAdd(sched, 1);
    __syncthreads();
    const int tile = s_item;
    __syncthreads();
    if(tile >= n_tiles) break;
    const int e = tile * BLOCK + tid;
    const bool active = e < a.n;
    const size_t env = active ? (size_t)e : 0;
    ParamsCompiledT<FOLLOW, false, FOLLOW> p = stage_dynamics_compiled<FOLLOW, false, FOLLOW>(sm_dyn, a.params, n, env, a.row);
    EnvState<Spec> st;
    load_state(st, a.state + env, n);
    DynInvariants d;
    {
        ParamsRW pg{a.params + env, n};
        dyn_invariants(d, pg, st);
    }
    float* hist_ptr = a.state + (size_t)S_HIST * n + env;
    uint64_t rng = a.rng[env];
    int ep_step = a.episode_step[env]; float ep_ret = a.episode_return[env]; bool truncated = a.truncated[env] != 0;
    uint64_t rng_at_reset = 0; bool reset_seen = false;
    Ring rg{oa.position[env], oa.current_episode_start[env], oa.full[env] != 0};
    float* ring = oa.replay + env * (size_t)oa.capacity * D;
    int* es = oa.episode_start + env * (size_t)oa.capacity;
    const int rows_valid = min(32, a.n - (tile * BLOCK + warp * 32));

    for(int t = 0; t < a.T; t++){
        if(truncated && active){                          // prologue_per_env (operations_generic_per_env.h:20-57)
            truncated = false; ep_step = 0; ep_ret = 0.0f;
            if(oa.sample_parameters){
                ParamsOverlay o;                          // sampled in registers: no dependent HBM round trips on the reset path
                o.init(a.row);
                if constexpr(DR && FOLLOW){ rng_at_reset = rng; reset_seen = true; }   // deferred parameter write-back, as in k_collect_ts
                if(!sample_parameters<DR, Spec::RNG_OOL, B200L2F_FAST_RESET != 0>(o, rng)) atomicExch(a.error_flag, 1);
                if constexpr(!FOLLOW) o.template flush<true>(ParamsRW{a.params + env, n});
                compile_dynamics_block<true, B200L2F_FAST_RESET != 0>(dyn_block_of_thread(sm_dyn), [&](int i){ return o[i]; });   // this thread's block only
                sample_state<Spec, ParamsOverlay, true, B200L2F_FAST_RESET != 0>(st, o, rng, hist_ptr, n);
                dyn_invariants<Spec, ParamsOverlay, B200L2F_FAST_RESET != 0>(d, o, st);
            }
            else{
                ParamsRW pg{a.params + env, n};
                sample_state<Spec, ParamsRW, true, B200L2F_FAST_RESET != 0>(st, pg, rng, hist_ptr, n);
                dyn_invariants(d, pg, st);
            }
            if(rg.full || rg.position > 0){
                const int previous = rg.position == 0 ? oa.capacity - 1 : rg.position - 1;
                ring[(size_t)previous * D + D - 1] = 1.0f;
                rg.current_start = rg.position;
            }
        }
        float obs[IN];
        observe_regs<Spec, true, true>(st, p, rng, obs);
        float o8[OUT], act[4];
        mlp_forward_ts<IN, OUT>(c, obs, o8);
        squash_sample<Spec::RNG_OOL, true>(o8, rng, act);  // interlude: evaluate_step in Mode<Rollout>
        RewardInputs ri;                                  // epilogue_per_env (:60-110)
        reward_inputs(ri, st);
        env_step_compiled<Spec, true, true, true, AXIAL>(st, p, d, act, rng, hist_ptr, n);
        const bool term = env_terminated(p, st.x);
        const float r = env_reward<true>(p, ri, act, st.x, term, d.dt);
        ep_ret += r; ep_step += 1;
        truncated = term || ep_step == a.step_limit;
        float* row = ring + (size_t)rg.position * D;
        // phase 1: obs | action | reward
        __syncwarp();
#pragma unroll
        for(int i = 0; i < IN; i++) win[lane * 33 + i] = obs[i];
#pragma unroll
        for(int i = 0; i < 4; i++) win[lane * 33 + IN + i] = act[i];
        win[lane * 33 + IN + 4] = r;
        stream_phase<IN + 5>(win, row, lane, rows_valid);
        // phase 2: next_obs | terminated | truncated
        observe_regs<Spec, true, true>(st, p, rng, obs);
#pragma unroll
        for(int i = 0; i < IN; i++) win[lane * 33 + i] = obs[i];
        win[lane * 33 + IN] = term ? 1.0f : 0.0f;
        win[lane * 33 + IN + 1] = truncated ? 1.0f : 0.0f;
        stream_phase<IN + 2>(win, row + IN + 5, lane, rows_valid);
        if(active) es[rg.position] = rg.current_start;
        ring_advance(rg, oa.capacity, truncated);
    }
    if(active){
        store_state(st, a.state + env, n);
        a.rng[env] = rng;
        a.episode_step[env] = ep_step; a.episode_return[env] = ep_ret; a.truncated[env] = truncated ? 1 : 0;
        oa.position[env] = rg.position; oa.current_episode_start[env] = rg.current_start; oa.full[env] = rg.full ? 1 : 0;
    }
    if constexpr(DR && FOLLOW){                           // the last reset's parameters reach the column here (k_collect_ts has the argument)
        if(active && reset_seen){
            ParamsOverlay o;
            o.init(a.row);
            uint64_t r = rng_at_reset;
            sample_parameters<DR, Spec::RNG_OOL, B200L2F_FAST_RESET != 0>(o, r);
            o.template flush<false>(ParamsRW{a.params + env, n});
        }
    }
    }   // tile loop
    mlp_ts_epilogue(c);
}

}  // namespace b200l2f
