// raptor_b200/csrc/step_only.cu -- b200l2f_step_repeated: T x rl_tools::step with ONE constant action, state in registers, no observe / policy / reward.
// This is the loop of the reference's own GPU benchmark (RT/src/rl/environments/l2f/cuda/benchmark.cu:98-128: simulate_parallel -- action 0, step(),
// state = next_state, N_ITERATIONS times) on this engine's data layout: per-environment parameters (the reference shares one parameter set per block),
// compiled dynamics block in shared memory, packed RK4.  It exists for the like-for-like line against that kernel (tools/ref_gpu_benchmark.sh) and as a
// fast-forward primitive ("advance every environment T control steps under a held action").
#include "launch.h"
#include "rollout_tc.cuh"

namespace b200l2f {
namespace {
template <class Spec, bool AXIAL>
__global__ void __launch_bounds__(BLOCK, 3) k_step_repeated(const float* __restrict__ params, float* __restrict__ state, uint64_t* __restrict__ rngs, int n, int T,
                                                             float a0, float a1, float a2, float a3, const __grid_constant__ RolloutArgs consts){
    extern __shared__ __align__(16) float sm_dyn[];
    constexpr int STRIDE = AXIAL ? C_DIM_AXIAL : C_DIM;
    const int e = blockIdx.x * BLOCK + threadIdx.x;
    const bool active = e < n;
    const size_t env = active ? (size_t)e : 0, nn = (size_t)n;
    const ParamsCompiledT<true> p = stage_dynamics_compiled<true, true, false, STRIDE>(sm_dyn, params, nn, env, consts.row0);
    EnvState<Spec> st;
    load_state(st, state + env, nn);
    DynInvariants d;
    {
        ParamsGlobal pg{params + env, nn};
        dyn_invariants(d, pg, st);
    }
    float* hist_ptr = state + (size_t)S_HIST * nn + env;
    uint64_t rng = rngs[env];
    const float action[4] = {a0, a1, a2, a3};
    if(Spec::H == 1 || active){
#pragma unroll 1
        for(int t = 0; t < T; t++) env_step_compiled<Spec, false, false, true, AXIAL>(st, p, d, action, rng, hist_ptr, nn);
    }
    if(!active) return;
    store_state(st, state + env, nn);
    rngs[env] = rng;
}
template <class Spec, bool AXIAL>
int launch(b200l2f_handle* h, int slot, const float* a, int T){
    auto kern = k_step_repeated<Spec, AXIAL>;
    const int smem = (AXIAL ? C_DIM_AXIAL : C_DIM) * BLOCK * (int)sizeof(float);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    RolloutArgs consts{};
    std::memcpy(consts.row0, h->row0, sizeof(consts.row0));
    kern<<<grid_for(h->n, BLOCK), BLOCK, smem, h->stream>>>(h->d_params, h->d_state[slot], h->d_rng, h->n, T, a[0], a[1], a[2], a[3], consts);
    h->last_kernel = "k_step_repeated";
    LAUNCH_CHECK();
    return B200L2F_OK;
}
}  // namespace
}  // namespace b200l2f

using namespace b200l2f;

extern "C" int b200l2f_step_repeated(b200l2f_handle* h, int slot, const float* action4, int32_t n_steps){
    if(!h) return fail(h, B200L2F_ERR_ARGUMENT, "step_repeated: null handle");
    CU(cudaSetDevice(h->cfg.device));
    if(!action4 || n_steps < 0) return fail(h, B200L2F_ERR_ARGUMENT, "step_repeated: bad arguments");
    int rc;
    if((rc = check_slot(h, slot)) || (rc = refresh_features(h))) return rc;
    if(h->features & 3) return fail(h, B200L2F_ERR_UNSUPPORTED, "step_repeated: needs noise-free, uniform MDP constants (the action-noise / Langevin constants ride in the launch's constant bank)");
    const bool axial = (h->features & 4) == 0;
    return dispatch_spec(h, [&](auto spec) -> int {
        using Spec = decltype(spec);
        return axial ? launch<Spec, true>(h, slot, action4, n_steps) : launch<Spec, false>(h, slot, action4, n_steps);
    });
}
