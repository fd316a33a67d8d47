// raptor_b200/csrc/rollout_ts.cu -- instantiations of k_rollout_raptor_ts (rollout_tc.cuh): THE default hot path (tcgen05, A operand and
// GRU hidden state in TMEM, persistent (tile, time-chunk) work queue).
#include "launch.h"
#include "rollout_tc.cuh"

namespace b200l2f {
namespace {
template <class Spec, bool FAST, bool UNIFORM, bool AXIAL, bool NOISE = false, int CTAS = 3, bool RECORD = true>
int launch_rollout_ts(b200l2f_handle* h, RolloutArgs a){
    auto kern = k_rollout_raptor_ts<Spec, FAST, UNIFORM, AXIAL, NOISE, CTAS, RECORD>;
    using TsSmem = TsSmemT<AXIAL, CTAS>;
    static bool configured[8] = {}; static int capacity[8] = {};
    int dev = h->cfg.device & 7;
    if(!configured[dev]){
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TsSmem::TOTAL));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, TsSmem::TOTAL));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if(std::getenv("B200L2F_VERBOSE")) std::fprintf(stderr, "[b200l2f] k_rollout_raptor_ts: occupancy API reports %d CTAs/SM on %d SMs\n", per_sm, sms);
        if(per_sm < TsSmem::CTAS) per_sm = TsSmem::CTAS;   // design points: 68 KB smem + 168 registers -> 3 CTAs/SM, 55 KB + 128 registers -> 4 (128 TMEM columns per CTA); a larger grid is harmless
        capacity[dev] = per_sm * sms;        // co-resident CTAs (3 x 148 = 444 on B200)
        configured[dev] = true;
    }
    int grid = 0, rc;
    if((rc = prepare_schedule(h, a, capacity[dev], &grid))) return rc;
    kern<<<grid, BLOCK, TsSmem::TOTAL, h->stream>>>(a, h->d_ts_image);
    h->last_kernel = CTAS == 4 ? "k_rollout_raptor_ts<CTAS=4>" : "k_rollout_raptor_ts<CTAS=3>";   // (the RECORD = false twin runs the same arithmetic: one name, one counters entry)
    LAUNCH_CHECK();
    return B200L2F_OK;
}
}  // namespace

int launch_raptor_ts(b200l2f_handle* h, const RolloutArgs& a, bool uniform, bool axial, bool noise){
    auto go = [&](auto spec) -> int {
        using Spec = decltype(spec);
        if(noise){
            if(!uniform) return fail(h, B200L2F_ERR_UNSUPPORTED, "rollout: the tcgen05 noise variant needs uniform MDP constants");
            return axial ? launch_rollout_ts<Spec, true, true, true, true>(h, a) : launch_rollout_ts<Spec, true, true, false, true>(h, a);
        }
        if(!uniform) return launch_rollout_ts<Spec, true, false, false>(h, a);
        if(!axial) return launch_rollout_ts<Spec, true, true, false>(h, a);
        // three or four resident CTAs per SM (see the kernel's comment): four once the tiles fill 4 x SMs slots; B200L2F_TS_CTAS=3|4 overrides
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
        const char* e = std::getenv("B200L2F_TS_CTAS");
        const int n_tiles = (a.n + BLOCK - 1) / BLOCK;
        const bool four = e ? e[0] == '4' : n_tiles >= 4 * sms;
        // no per-step output requested (evaluation / benchmark launches): the instantiation without the recording branches; B200L2F_TS_RECORD=1 forces the other
        static const bool force_record = [](){ const char* r = std::getenv("B200L2F_TS_RECORD"); return r && r[0] == '1'; }();
        const bool record = force_record || a.out_states || a.out_obs || a.out_actions || a.out_rewards || a.out_term;
        if(record) return four ? launch_rollout_ts<Spec, true, true, true, false, 4, true>(h, a) : launch_rollout_ts<Spec, true, true, true, false, 3, true>(h, a);
        return four ? launch_rollout_ts<Spec, true, true, true, false, 4, false>(h, a) : launch_rollout_ts<Spec, true, true, true, false, 3, false>(h, a);
    };
    return h->kind == KIND_DEFAULT ? go(SpecDefault{}) : go(SpecRaptor{});
}
}  // namespace b200l2f
