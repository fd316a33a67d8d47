// raptor_b200/csrc/collect.cu -- instantiations of k_collect (mlp.cuh): PPO collection with the actor on fp32 CUDA cores.
#include "launch.h"

namespace b200l2f {

int launch_collect_fp32(b200l2f_handle* h, const CollectArgs& a){
    auto go = [&](auto spec, auto dr_c) -> int {
        using Spec = SpecCompactCode<decltype(spec)>;
        constexpr bool DR = decltype(dr_c)::value;
        constexpr int IN = Spec::OBS_DIM;
        auto kern = k_collect<Spec, DR>;
        const size_t smem = sizeof(float) * (MlpImg<IN, 4>::SIZE + (size_t)P_DYN_DIM * BLOCK + (size_t)CollectSlab<IN>::FLOATS * (BLOCK / 32));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid_for(a.n, BLOCK), BLOCK, smem, h->stream>>>(a);
        h->last_kernel = "k_collect";
        LAUNCH_CHECK();
        return (int)B200L2F_OK;
    };
    // DEFAULT spec (H = 16, OBS 82: the PPO zoo's environment, rl/zoo/l2f/ppo.h): 193 KB of shared memory, one CTA per SM
    if(h->kind == KIND_DEFAULT) return h->dr ? go(SpecDefault{}, std::true_type{}) : go(SpecDefault{}, std::false_type{});
    if(h->kind == KIND_RAPTOR) return h->dr ? go(SpecRaptor{}, std::true_type{}) : go(SpecRaptor{}, std::false_type{});
    return h->dr ? go(SpecTeacher{}, std::true_type{}) : go(SpecTeacher{}, std::false_type{});
}
}  // namespace b200l2f
