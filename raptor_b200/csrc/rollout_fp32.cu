// raptor_b200/csrc/rollout_fp32.cu -- instantiations of k_rollout_raptor (kernels.cuh): the fused rollout with the actor on fp32 CUDA cores.
#include "launch.h"

namespace b200l2f {
namespace {
template <class Spec, bool NOISE, bool FAST, bool CONSTW, bool ROLLED = false>
int launch_rollout_raptor(b200l2f_handle* h, const RolloutArgs& a){
    constexpr int IN = 22, HD = 16, OUT = 4;
    constexpr int IMG = RaptorImage<IN, HD, OUT>::SIZE;
    auto kern = k_rollout_raptor<Spec, IN, HD, OUT, NOISE, FAST, CONSTW, ROLLED>;
    const size_t smem = (size_t)((CONSTW ? 0 : IMG) + P_DYN_DIM * BLOCK + (ROLLED ? RaptorScratch<IN, HD>::ROWS * BLOCK : 0)) * sizeof(float);
    static bool configured[8] = {};   // per device
    int dev = h->cfg.device & 7;
    if(!configured[dev]){
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    if constexpr(CONSTW){
        kern<<<grid_for(a.n, BLOCK), BLOCK, smem, h->stream>>>(a, *reinterpret_cast<const WeightBlock<IMG>*>(h->h_image.data()));
    }
    else{
        kern<<<grid_for(a.n, BLOCK), BLOCK, smem, h->stream>>>(a, WeightBlock<1>{});
    }
    h->last_kernel = "k_rollout_raptor";
    LAUNCH_CHECK();
    return B200L2F_OK;
}
}  // namespace

int launch_raptor_fp32(b200l2f_handle* h, const RolloutArgs& a, bool noise, bool fast, bool constw, bool rolled){
    auto go = [&](auto spec) -> int {
        using Spec = decltype(spec);
        if(noise) return fast ? launch_rollout_raptor<Spec, true, true, true>(h, a) : launch_rollout_raptor<Spec, true, false, true>(h, a);
        if(!fast) return launch_rollout_raptor<Spec, false, false, true>(h, a);
        if(constw) return launch_rollout_raptor<Spec, false, true, true>(h, a);
        return rolled ? launch_rollout_raptor<Spec, false, true, false, true>(h, a) : launch_rollout_raptor<Spec, false, true, false, false>(h, a);
    };
    return h->kind == KIND_DEFAULT ? go(SpecDefault{}) : go(SpecRaptor{});
}
}  // namespace b200l2f
