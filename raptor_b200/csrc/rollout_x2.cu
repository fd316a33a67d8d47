// raptor_b200/csrc/rollout_x2.cu -- instantiation of k_rollout_raptor_x2 (rollout_x2.cuh): two environments per thread on the packed fp32 pipe.
#include "launch.h"
#include "rollout_x2.cuh"

namespace b200l2f {

int launch_raptor_x2(b200l2f_handle* h, const RolloutArgs& a_in){
    RolloutArgs a = a_in;
    auto kern = k_rollout_raptor_x2<SpecRaptor>;
    static bool configured[8] = {}; static int capacity[8] = {};
    const int dev = h->cfg.device & 7;
    if(!configured[dev]){
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X2Smem::TOTAL));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, X2Smem::TOTAL));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        if(std::getenv("B200L2F_VERBOSE")) std::fprintf(stderr, "[b200l2f] k_rollout_raptor_x2: occupancy API reports %d CTAs/SM on %d SMs\n", per_sm, sms);
        if(per_sm < 1) return fail(h, B200L2F_ERR_CUDA, "k_rollout_raptor_x2 does not fit on this device");
        // design point: 256 registers x 128 threads, 104 KB of shared memory and 256 TMEM columns per CTA -> exactly two CTAs per SM.  The occupancy API
        // answers before the shared-memory carve-out preference takes effect (it reported 1 here and 1 for the TS kernel's 3); a larger grid is harmless
        per_sm = 2;
        capacity[dev] = per_sm * sms;
        configured[dev] = true;
    }
    int grid = 0, rc;
    if((rc = prepare_schedule(h, a, capacity[dev], &grid, X2Smem::ENVS))) return rc;
    kern<<<grid, BLOCK, X2Smem::TOTAL, h->stream>>>(a, h->d_ts_image);
    h->last_kernel = "k_rollout_raptor_x2";
    LAUNCH_CHECK();
    return B200L2F_OK;
}

}  // namespace b200l2f
