// raptor_b200/csrc/collect_ts.cu -- instantiations of k_collect_ts (mlp_tc.cuh): PPO collection with the actor on tcgen05.
#include <cstdlib>
#include "launch.h"
#include "mlp_tc.cuh"
#include "collect_lag.cuh"

namespace b200l2f {

int launch_collect_ts(b200l2f_handle* h, const CollectArgs& a, bool follow, bool row_axial){
    auto gots2 = [&](auto spec, auto dr_c, auto follow_c, auto axial_c) -> int {
        using Spec = SpecCompactCode<decltype(spec)>;
        constexpr bool DR = decltype(dr_c)::value;
        using SM = MlpTsSmem<Spec::OBS_DIM, 4>;
        // default: k_collect_ts (resets inline); B200L2F_COLLECT_LAG=1 selects k_collect_lag (resets on a fifth warp, collect_lag.cuh: bit-identical datasets,
        // measured 10 % slower on config 4 -- profiles/r02_exp5_collect_reset_warp.log -- kept as the starting point for a register-rebalanced version)
        static const bool lag = [](){ const char* e = std::getenv("B200L2F_COLLECT_LAG"); return e && e[0] == '1'; }();
        auto kern = lag ? k_collect_lag<Spec, DR, decltype(follow_c)::value, decltype(axial_c)::value> : k_collect_ts<Spec, DR, decltype(follow_c)::value, decltype(axial_c)::value>;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL_COLLECT));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if(!h->d_sched){ CU(cudaMalloc(&h->d_sched, sizeof(int) * 64)); h->sched_ints = 64; }
        CU(cudaMemsetAsync(h->d_sched, 0, sizeof(int), h->stream));
        const int n_tiles = grid_for(a.n, BLOCK);
        const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;    // ~110 KB smem, 256 TMEM columns per CTA -> 2 CTAs/SM, persistent tile loop
        kern<<<grid, lag ? LAG_THREADS : BLOCK, SM::TOTAL_COLLECT, h->stream>>>(a, h->d_mlp_tc_image, h->d_sched);
        h->last_kernel = lag ? "k_collect_lag" : "k_collect_ts";
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    };
    auto gots = [&](auto spec, auto dr_c) -> int {
        if(follow) return row_axial ? gots2(spec, dr_c, std::true_type{}, std::true_type{}) : gots2(spec, dr_c, std::true_type{}, std::false_type{});
        return gots2(spec, dr_c, std::false_type{}, std::false_type{});
    };
    if(h->kind == KIND_DEFAULT) return launch_collect_ts_default(h, a, follow, row_axial);   // collect_ts_default.cu (its own translation unit: compile time)
    if(h->kind == KIND_RAPTOR) return h->dr ? gots(SpecRaptor{}, std::true_type{}) : gots(SpecRaptor{}, std::false_type{});
    return h->dr ? gots(SpecTeacher{}, std::true_type{}) : gots(SpecTeacher{}, std::false_type{});
}
}  // namespace b200l2f
