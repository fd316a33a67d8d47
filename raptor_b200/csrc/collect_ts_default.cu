// raptor_b200/csrc/collect_ts_default.cu -- k_collect_ts (mlp_tc.cuh) for the DEFAULT specification: the environment of the reference's PPO zoo
// (INC/rl/zoo/l2f/ppo.h:28-33; L2F/parameters/default.h:159: H = 16 action history, 82-wide observation, 97-float dataset rows).  The first dense
// layer is a K = 88 tensor-core operand (eleven K = 8 blocks per plane, TMEM plan "wide" in mlp_tc.cuh); the 87 KB weight image, the dynamics block
// the per-warp write-back windows (which double as the observation scratch) and the action-history rings take 207 KB of shared memory: one CTA per SM,
// persistent (tile, time-chunk) queue.
#include <cstdlib>
#include "launch.h"
#include "mlp_tc.cuh"

namespace b200l2f {

int launch_collect_ts_default(b200l2f_handle* h, const CollectArgs& a, bool follow, bool row_axial){
    auto go = [&](auto dr_c, auto follow_c, auto axial_c) -> int {
        using Spec = SpecCompactCode<SpecDefault>;
        using SM = MlpTsSmem<Spec::OBS_DIM, 4>;
        auto kern = k_collect_ts<Spec, decltype(dr_c)::value, decltype(follow_c)::value, decltype(axial_c)::value>;
        constexpr int SMEM = SM::TOTAL_COLLECT + SM::HIST_RING;   // + the CTA's action-history rings
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        const int n_tiles = grid_for(a.n, BLOCK);
        const int grid = n_tiles < sms ? n_tiles : sms;
        // one CTA per SM: a tile count that is not a multiple of the grid costs a whole extra round (512 tiles on 148 SMs: 4 rounds for 3.46 of work).  Cutting the
        // collection into time chunks (k_collect_ts' chunk-major queue) evens that out.  This translation unit is compiled with -Xptxas -dlcm=cg (build.py: every global
        // load through L2), which is what makes the cross-SM hand-over of a tile between chunks coherent.  B200L2F_COLLECT_CHUNKS overrides (1 = off).
        CollectArgs b = a;
        int chunks = 1;
        if(n_tiles > grid && n_tiles % grid != 0){          // rounds of the launch in tile-times: ceil(tiles * c / grid) / c, plus ~2 % per extra chunk (measured: hand-over + re-staging)
            double best = 1e30;
            for(int c = 1; c <= 4; c++){
                const double cost = (double)((n_tiles * c + grid - 1) / grid) / c * (1.0 + 0.02 * (c - 1));
                if(cost < best - 1e-9){ best = cost; chunks = c; }
            }
        }
        if(const char* e = std::getenv("B200L2F_COLLECT_CHUNKS")) chunks = std::atoi(e) > 0 ? std::atoi(e) : 1;
        if(chunks > a.T) chunks = a.T > 0 ? a.T : 1;
        b.n_chunks = chunks; b.chunk_steps = chunks > 1 ? (a.T + chunks - 1) / chunks : a.T + 1;
        if(chunks > 1 && (b.n_chunks - 1) * b.chunk_steps >= a.T) b.n_chunks = (a.T + b.chunk_steps - 1) / b.chunk_steps;   // no empty chunks
        const int ints = 1 + n_tiles;
        if(!h->d_sched || h->sched_ints < ints){ if(h->d_sched) CU(cudaFree(h->d_sched)); CU(cudaMalloc(&h->d_sched, sizeof(int) * ints)); h->sched_ints = ints; }
        CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int) * ints, h->stream));
        kern<<<grid, BLOCK, SMEM, h->stream>>>(b, h->d_mlp_tc_image, h->d_sched);
        h->last_kernel = "k_collect_ts<DEFAULT>";
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    };
    auto by_follow = [&](auto dr_c) -> int {
        if(follow) return row_axial ? go(dr_c, std::true_type{}, std::true_type{}) : go(dr_c, std::true_type{}, std::false_type{});
        return go(dr_c, std::false_type{}, std::false_type{});
    };
    return h->dr ? by_follow(std::true_type{}) : by_follow(std::false_type{});
}
}  // namespace b200l2f
