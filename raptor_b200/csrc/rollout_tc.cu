// raptor_b200/csrc/rollout_tc.cu -- instantiations of k_rollout_raptor_tc (rollout_tc.cuh): tcgen05 actor GEMMs with the A operand in shared memory.
#include "launch.h"
#include "rollout_tc.cuh"

namespace b200l2f {
namespace {
template <class Spec, bool FAST, bool UNIFORM, bool G1_TC>
int launch_rollout_tc(b200l2f_handle* h, const RolloutArgs& a){
    auto kern = k_rollout_raptor_tc<Spec, FAST, UNIFORM, G1_TC>;
    static bool configured[8] = {};
    int dev = h->cfg.device & 7;
    if(!configured[dev]){
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcSmem::TOTAL));
        configured[dev] = true;
    }
    kern<<<grid_for(a.n, BLOCK), BLOCK, TcSmem::TOTAL, h->stream>>>(a, h->d_tc_image);
    h->last_kernel = "k_rollout_raptor_tc";
    LAUNCH_CHECK();
    return B200L2F_OK;
}
}  // namespace

int launch_raptor_tc(b200l2f_handle* h, const RolloutArgs& a, bool fast, bool uniform, bool g1_tc){
    auto go = [&](auto spec) -> int {
        using Spec = decltype(spec);
        if(!fast) return launch_rollout_tc<Spec, false, false, true>(h, a);
        if(!uniform) return launch_rollout_tc<Spec, true, false, true>(h, a);
        return g1_tc ? launch_rollout_tc<Spec, true, true, true>(h, a) : launch_rollout_tc<Spec, true, true, false>(h, a);
    };
    return h->kind == KIND_DEFAULT ? go(SpecDefault{}) : go(SpecRaptor{});
}
}  // namespace b200l2f
