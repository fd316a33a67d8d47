// raptor_b200/csrc/collect_lag.cuh -- k_collect_lag: PPO collection (rl_tools::collect, same dataset contract as k_collect_ts in mlp_tc.cuh) with the
// environment RESETS taken off the tile's critical path.
//
// Why: a reset (sample_initial_parameters + sample_initial_state, ~1600 dependent instructions, ~40 sequential RNG draws) is divergent work: in
// k_collect_ts the warp of a terminating environment executes it while its other 31 lanes idle and the CTA's other three warps wait at the next MMA
// barrier.  With a random-init actor 3.7 % of the rows end an episode, so 70 % of the warp-steps (99 % of the CTA-steps) pay one full reset latency:
// half of the kernel's time (measured: 10.0 ms without terminations, 19.6 ms with, 262 144 envs x 256 steps).
//
// How: the CTA gets a fifth warp that does nothing but resets.  A lane whose episode ended parks its state in HBM, raises a flag and SITS OUT (its
// row of the MMAs computes garbage that nobody reads) until the reset warp has re-sampled its parameters and state -- normally one iteration.  Every
// lane therefore carries its own step counter: an iteration of the CTA advances the running lanes by one step each, and the tile ends when all 128
// lanes have written their T + 1 rows (T + number of sit-out iterations of the slowest lane).  Each environment still sees exactly the reference's
// sequence -- its own RNG stream, reset before the next observation (operations_generic_per_env.h:17-25) -- so the dataset is bit-identical to
// k_collect_ts's; only the time at which a row is produced moves.  The reset warp packs the pending requests of all four warps into its 32 lanes
// (dense SIMT instead of 1-2 active lanes per divergent warp).
//
// Hand-over: through the sitting-out lane's own row of the write-back window (37+ floats of shared memory nobody else touches while the lane
// contributes no row): the owner leaves its RNG state there, the reset warp returns the 31 words the owner cannot derive itself (integrated state,
// disturbances, trajectory type, action history, RNG state, mass).  No HBM round trip and no device-scope fence on the tile's critical path.
//
// Rows: lanes of a warp are at different steps, so the warp's 32 rows are no longer one contiguous run; a warp whose lanes agree (always until
// its first termination) still sends them as one TMA bulk copy, otherwise row by row (two coalesced store instructions per row).
#pragma once
#include "mlp_tc.cuh"

namespace b200l2f {

constexpr int LAG_THREADS = BLOCK + 32;      // four tile warps + the reset warp
enum { LAG_IDLE = 0, LAG_REQUESTED = 1, LAG_DONE = 2 };
enum { MB_X = 0, MB_FORCE = 17, MB_TORQUE = 20, MB_TRAJ = 23, MB_HIST = 24, MB_RNG = 28, MB_MASS = 30, MB_WORDS = 31 };   // mailbox words in the owner's window row

struct LagShared {
    int flag[BLOCK];        // per slot: LAG_IDLE / LAG_REQUESTED (owner -> reset warp) / LAG_DONE (reset warp -> owner)
    int done;               // lanes of the current tile that have written all their rows
    int tile_over;          // tile index + 1 once the tile warps have left the tile (releases the reset warp)
    int item;               // tile scheduler
};

template <class Spec, bool DR, bool FOLLOW, bool AXIAL>
__global__ void __launch_bounds__(LAG_THREADS, 2) k_collect_lag(const __grid_constant__ CollectArgs a, const float* __restrict__ tc_image, int* __restrict__ sched){
    constexpr int IN = Spec::OBS_DIM, OUT = 4;
    constexpr int D = IN + 15, W = IN + 12;
    using SM = MlpTsSmem<IN, OUT>;
    using I = MlpTcImage<IN, OUT>;
    extern __shared__ __align__(1024) unsigned char smraw[];
    __shared__ LagShared sh;
    float* sm_dyn = reinterpret_cast<float*>(smraw + SM::DYN);
    if(threadIdx.x == 0) sh.tile_over = 0;
    TsCtx c = mlp_ts_prologue<IN, OUT>(smraw, tc_image);     // all 160 threads (barriers inside)
    c.probe = &sh.done; c.probe_value = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tc::uniform_warp_index();
    const bool reset_warp = warp == BLOCK / 32;
    const size_t n = (size_t)a.n;
    const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
    float* slab = reinterpret_cast<float*>(smraw + SM::SLAB) + (size_t)(warp & 3) * 32 * D;   // tile warps: image of the warp's 32 dataset rows
    float* myrow = slab + lane * D;
    if(!reset_warp){
#pragma unroll
        for(int i = W; i < D; i++) myrow[i] = 0.0f;          // the learner's columns leave as zeros
    }
    bool store_pending = false;

    for(;;){
    if(tid == 0) sh.item = atomicAdd(sched, 1);
    if(tid < BLOCK) sh.flag[tid] = LAG_IDLE;
    if(tid == 0) sh.done = 0;
    __syncthreads();
    const int tile = sh.item;
    __syncthreads();
    if(tile >= n_tiles) break;

    if(reset_warp){
        // ---------------------------------------------------------------- the reset warp: serve requests until the tile is over
        uint32_t idle_spins = 0;
        for(;;){
            if(*(volatile int*)&sh.tile_over == tile + 1) break;
            uint32_t m[4]; int total = 0;
#pragma unroll
            for(int g = 0; g < 4; g++){ m[g] = __ballot_sync(0xffffffffu, *(volatile int*)&sh.flag[32 * g + lane] == LAG_REQUESTED); total += __popc(m[g]); }
            if(total == 0){
                __nanosleep(128);
                if(++idle_spins > (1u << 26)) __trap();      // a lost tile: report instead of hanging the GPU
                continue;
            }
            idle_spins = 0;
            int slot = -1, k = lane;                         // lane j takes the j-th pending request of the CTA
#pragma unroll
            for(int g = 0; g < 4; g++){
                const int cnt = __popc(m[g]);
                if(slot < 0 && k >= 0){
                    if(k < cnt){ slot = 32 * g + (int)__fns(m[g], 0, k + 1); k = -1; }
                    else k -= cnt;
                }
            }
            if(slot >= 0){
                __threadfence_block();                       // the owner's request word (written before it raised the flag)
                const size_t env = (size_t)tile * BLOCK + slot;
                float* mb = reinterpret_cast<float*>(smraw + SM::SLAB) + (size_t)slot * D;   // the owner's window row
                uint64_t rng = (uint64_t)__float_as_uint(mb[MB_RNG]) | ((uint64_t)__float_as_uint(mb[MB_RNG + 1]) << 32);
                float* hist_ptr = a.state + (size_t)S_HIST * n + env;
                ParamsOverlay o;                             // sampled in registers: no dependent HBM round trips
                o.init(a.row);
                if(!sample_parameters<DR, Spec::RNG_OOL, B200L2F_FAST_RESET != 0>(o, rng)) atomicExch(a.error_flag, 1);
                if constexpr(DR || !FOLLOW) o.template flush<!FOLLOW>(ParamsRW{a.params + env, n});
                compile_dynamics_block<true, B200L2F_FAST_RESET != 0>(sm_dyn + (size_t)slot * C_DIM, [&](int i){ return o[i]; });   // the owner's block (it sits out: nobody reads it now)
                EnvState<Spec> st;                           // built from scratch; the dead Langevin target stays with the owner (see the resume path)
                sample_state<Spec, ParamsOverlay, true, B200L2F_FAST_RESET != 0>(st, o, rng, hist_ptr, n);
#pragma unroll
                for(int i = 0; i < X_DIM; i++) mb[MB_X + i] = st.x[i];
#pragma unroll
                for(int i = 0; i < 3; i++){ mb[MB_FORCE + i] = st.force[i]; mb[MB_TORQUE + i] = st.torque[i]; }
                mb[MB_TRAJ] = (float)st.traj_type;
                if constexpr(Spec::H == 1){
#pragma unroll
                    for(int i = 0; i < 4; i++) mb[MB_HIST + i] = st.hist[i];
                }
                mb[MB_RNG] = __uint_as_float((uint32_t)rng); mb[MB_RNG + 1] = __uint_as_float((uint32_t)(rng >> 32));
                mb[MB_MASS] = o[P_MASS];
                if constexpr(!FOLLOW || Spec::H != 1) __threadfence();   // the owner reads parameter columns / the action ring from HBM
                else __threadfence_block();
                *(volatile int*)&sh.flag[slot] = LAG_DONE;
            }
            __syncwarp();
        }
    }
    else{
        // ---------------------------------------------------------------- the four tile warps
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
        const int warp_env0 = tile * BLOCK + warp * 32;
        const int rows_valid = min(32, a.n - warp_env0);
        const bool bulk_ok = a.bulk_rows != 0 && rows_valid == 32;
        const int n_active = min(BLOCK, a.n - tile * BLOCK);
        int t = 0;                                            // this lane's step
        bool running = active, finished = !active;
        // park the state and ask for a reset (operations_generic_per_env.h:17-25: a truncated environment is re-sampled before its next step)
        auto request_reset = [&](){
            myrow[MB_RNG] = __uint_as_float((uint32_t)rng); myrow[MB_RNG + 1] = __uint_as_float((uint32_t)(rng >> 32));
            __threadfence_block();
            *(volatile int*)&sh.flag[tid] = LAG_REQUESTED;
            running = false;
        };
        // the request word goes into the window row: an earlier bulk copy of the window must have read it first
        auto request_if = [&](bool want){
            if(__any_sync(0xffffffffu, want) && store_pending){ if(lane == 0) tc::bulk_store_wait_read(); store_pending = false; __syncwarp(); }
            if(want) request_reset();
        };
        request_if(running && truncated && a.T > 0);

        for(uint32_t iteration = 0; ; iteration++){
            if(iteration > 64u * (uint32_t)(a.T + 2) + 4096u) __trap();   // a lost reset: report instead of hanging the GPU
            if(!running && !finished && *(volatile int*)&sh.flag[tid] == LAG_DONE){   // back from the reset warp
                if constexpr(!FOLLOW || Spec::H != 1) __threadfence();
                else __threadfence_block();
#pragma unroll
                for(int i = 0; i < X_DIM; i++) st.x[i] = myrow[MB_X + i];
#pragma unroll
                for(int i = 0; i < 3; i++){ st.force[i] = myrow[MB_FORCE + i]; st.torque[i] = myrow[MB_TORQUE + i]; }
#pragma unroll
                for(int i = 0; i < 4; i++) st.last_action[i] = 0.0f;
                st.current_step = 0;
                st.traj_type = (int)myrow[MB_TRAJ];
                if constexpr(Spec::LANGEVIN){                 // sample_state<.., KEEP_DEAD_TARGET>: a POSITION episode keeps the previous target (never read)
                    if(st.traj_type == 1){
#pragma unroll
                        for(int i = 0; i < 12; i++) st.lang[i] = 0.0f;
                    }
                }
                if constexpr(Spec::H == 1){
#pragma unroll
                    for(int i = 0; i < 4; i++) st.hist[i] = myrow[MB_HIST + i];
                }
                rng = (uint64_t)__float_as_uint(myrow[MB_RNG]) | ((uint64_t)__float_as_uint(myrow[MB_RNG + 1]) << 32);
                {
                    const float mass = myrow[MB_MASS];
                    auto pa = [&](int i) -> float {           // dyn_invariants' accessor: what the reset changed comes from the mailbox / the fresh dynamics block
                        if(i == P_MASS) return mass;
                        if(i == P_JINV) return p.c(C_JID); if(i == P_JINV + 4) return p.c(C_JID + 1); if(i == P_JINV + 8) return p.c(C_JID + 2);
                        return p[i];
                    };
                    struct Acc { decltype(pa)& f; __device__ __forceinline__ float operator[](int i) const { return f(i); } } acc{pa};
                    dyn_invariants<Spec, Acc, B200L2F_FAST_RESET != 0>(d, acc, st);
                }
                sh.flag[tid] = LAG_IDLE;
                truncated = false; ep_step = 0; ep_ret = 0.0f;
                running = true;
            }
            // the previous iteration's bulk store has read the window (in flight since then: the wait is free)
            if(store_pending){ if(lane == 0) tc::bulk_store_wait_read(); store_pending = false; }
            __syncwarp();
            const bool last = t == a.T;                       // this lane's final observation (operations_generic.h:122-129)
            const int row_t = running ? t : -1;               // the row this lane contributes in this iteration (none: its window row is the mailbox)
            float obs[Spec::H == 1 ? IN : 1];
            if constexpr(Spec::H == 1){
                observe_regs<Spec, true, true>(st, p, rng, obs);
                if(row_t >= 0){
#pragma unroll
                    for(int i = 0; i < IN; i++) myrow[i] = obs[i];
                }
            }
            else if(running) observe_to_scratch<Spec, true>(st, p, rng, hist_ptr, n, myrow, 1);   // reads the action ring in HBM
            const bool finishing = running && last;
            if(finishing){                                    // all rows written after this iteration: final bookkeeping now, the code below computes garbage for this lane
                store_state(st, a.state + env, n);
                a.rng[env] = rng;
                a.episode_step[env] = ep_step; a.episode_return[env] = ep_ret; a.truncated[env] = truncated ? 1 : 0;
                running = false; finished = true;
            }
            float mean[OUT], act[4];
            if constexpr(Spec::H == 1) mlp_forward_ts<IN, OUT, true>(c, obs, mean);
            else mlp_forward_ts_from<IN, OUT, true>(c, [&](int k){ return myrow[k]; }, mean);
            // uniform exit: `done` (read behind the MLP's first barrier, incremented at the end of an iteration, i.e. never between that barrier and
            // the next one) counts the lanes whose rows were all written in EARLIER iterations
            if(c.probe_value == n_active){
                break;
            }
            float vals[12];
            {
                float lp = 0.0f;
#pragma unroll
                for(int i = 0; i < 4; i++){                   // epilogue (operations_generic_per_env.h:43-58)
                    const float ls = c.sm_b[I::LOG_STD + i];
                    act[i] = rng_normal_t<Spec::RNG_OOL, true>(rng, mean[i], ex2_approx(ls * LOG2E));
                    lp += normal_log_prob(mean[i], ls, act[i]);
                }
                RewardInputs ri;
                reward_inputs(ri, st);
                if(Spec::H == 1 || running) env_step_compiled<Spec, B200L2F_COLLECT_ROLLED_RK4 != 0, true, true, AXIAL, false>(st, p, d, act, rng, hist_ptr, n);
                if(running) langevin_update_compiled<Spec, true>(st, p, rng, d.dt);   // a sitting-out lane keeps its (dead) Langevin target as the reset leaves it (sample_state<.., KEEP_DEAD_TARGET>)
                const bool term = env_terminated(p, st.x);
                const float r = env_reward<true>(p, ri, act, st.x, term, d.dt);
                ep_ret += r; ep_step += 1;
                truncated = term || (a.step_limit > 0 && ep_step >= a.step_limit);
                const bool full = row_t >= 0 && row_t < a.T;  // a step row; the final row carries the observation only
#pragma unroll
                for(int i = 0; i < 4; i++){ vals[i] = full ? mean[i] : 0.0f; vals[4 + i] = full ? act[i] : 0.0f; }
                vals[8] = full ? lp : 0.0f; vals[9] = full ? r : 0.0f; vals[10] = full && term ? 1.0f : 0.0f; vals[11] = full && truncated ? 1.0f : 0.0f;
            }
            if(row_t >= 0){
#pragma unroll
                for(int i = 0; i < 12; i++) myrow[IN + i] = vals[i];
            }
            // ---- write-back
            const int t0 = __shfl_sync(0xffffffffu, row_t, 0);
            const bool same = __all_sync(0xffffffffu, row_t == t0);
            if(same && t0 < 0){}                              // the whole warp sits out
            else if(same && bulk_ok){                   // the warp's 32 rows are one contiguous run: one bulk copy
                tc::fence_async_smem();
                __syncwarp();
                if(lane == 0) tc::bulk_store(a.dataset + ((size_t)t0 * n + warp_env0) * D, slab, 32 * D * 4);
                store_pending = true;
            }
            else{
                // row by row: every lane publishes where ITS row goes (0 = none), the warp copies row r with two predicated, coalesced stores
                __syncwarp();
                const unsigned long long mine = row_t >= 0 ? (unsigned long long)(a.dataset + ((size_t)row_t * n + env) * D) : 0ull;
#pragma unroll 8
                for(int r = 0; r < 32; r++){
                    float* dst = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, mine, r));
                    const float* src = slab + r * D;
                    if(dst){                                  // warp-uniform; no row, no stores
#pragma unroll
                        for(int cc = 0; cc < D; cc += 32) if(cc + lane < D) dst[cc + lane] = src[cc + lane];
                    }
                }
                __syncwarp();
            }
            // ---- this lane's next step
            if(running) t += 1;
            request_if(running && truncated && t < a.T);
            if(finishing) atomicAdd(&sh.done, 1);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");        // every tile warp has left the loop
        if(tid == 0) *(volatile int*)&sh.tile_over = tile + 1;
    }
    }   // tile loop
    if(store_pending && lane == 0) tc::bulk_store_wait_all();   // the window must outlive the copy
    mlp_ts_epilogue(c);
}

}  // namespace b200l2f
