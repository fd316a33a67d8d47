// raptor_b200/csrc/collect_ts_default.cu -- k_collect_ts (mlp_tc.cuh) for the DEFAULT specification: the environment of the reference's PPO zoo
// (INC/rl/zoo/l2f/ppo.h:28-33; L2F/parameters/default.h:159: H = 16 action history, 82-wide observation, 97-float dataset rows).  The first dense
// layer is a K = 88 tensor-core operand (eleven K = 8 blocks per plane, TMEM plan "wide" in mlp_tc.cuh); the 87 KB weight image, the dynamics block
// and the per-warp write-back windows (which double as the observation scratch) take 175 KB of shared memory: one CTA per SM, persistent tile loop.
#include "launch.h"
#include "mlp_tc.cuh"

namespace b200l2f {

int launch_collect_ts_default(b200l2f_handle* h, const CollectArgs& a, bool follow, bool row_axial){
    auto go = [&](auto dr_c, auto follow_c, auto axial_c) -> int {
        using Spec = SpecCompactCode<SpecDefault>;
        using SM = MlpTsSmem<Spec::OBS_DIM, 4>;
        auto kern = k_collect_ts<Spec, decltype(dr_c)::value, decltype(follow_c)::value, decltype(axial_c)::value>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL_COLLECT));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if(!h->d_sched){ CU(cudaMalloc(&h->d_sched, sizeof(int) * 64)); h->sched_ints = 64; }
        CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int), h->stream));
        const int n_tiles = grid_for(a.n, BLOCK);
        const int grid = n_tiles < sms ? n_tiles : sms;
        kern<<<grid, BLOCK, SM::TOTAL_COLLECT, h->stream>>>(a, h->d_mlp_tc_image, h->d_sched);
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
