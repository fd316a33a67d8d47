// raptor_b200/csrc/rollout_mlp.cu -- instantiations of k_rollout_mlp / k_mlp_step (mlp.cuh): MLP actors on fp32 CUDA cores.
#include "launch.h"

namespace b200l2f {

int launch_mlp_fp32(b200l2f_handle* h, const RolloutArgs& a){
        auto gomlp = [&](auto spec, auto out_c) -> int {
            using Spec = decltype(spec);
            constexpr int OUT = decltype(out_c)::value, IN = Spec::OBS_DIM;
            constexpr int ROWS = IN > MLP_HD ? IN : MLP_HD;
            auto kern = k_rollout_mlp<Spec, OUT>;
            const size_t smem = sizeof(float) * (MlpImg<IN, OUT>::SIZE + (size_t)P_DYN_DIM * BLOCK + (size_t)ROWS * BLOCK);
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid_for(a.n, BLOCK), BLOCK, smem, h->stream>>>(a, h->pol.standardize);
            h->last_kernel = "k_rollout_mlp";
            LAUNCH_CHECK();
            return (int)B200L2F_OK;
        };
        using O4 = std::integral_constant<int, 4>; using O8 = std::integral_constant<int, 8>;
        const bool o8 = h->pol.output_dim == 8;
    return dispatch_spec(h, [&](auto spec){ return o8 ? gomlp(spec, O8{}) : gomlp(spec, O4{}); });
}

int launch_mlp_step(b200l2f_handle* h, const float* d_obs, int ld, float* d_act){
    int rc;
        const int has_std = h->pol.standardize, has_ls = h->pol.head == B200L2F_HEAD_PPO_GAUSSIAN, head = h->pol.head;
        auto go = [&](auto in_c, auto out_c) -> int {
            constexpr int IN = decltype(in_c)::value, OUT = decltype(out_c)::value;
            constexpr int ROWS = IN > MLP_HD ? IN : MLP_HD;
            auto kern = k_mlp_step<IN, OUT>;
            const size_t smem = sizeof(float) * (MlpImg<IN, OUT>::SIZE + (size_t)ROWS * BLOCK);
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid_for(h->n, BLOCK), BLOCK, smem, h->stream>>>(h->d_blob, has_std, has_ls, head, (const float*)d_obs, ld, h->d_rng, (float*)d_act, h->n);
            LAUNCH_CHECK();
            return (int)B200L2F_OK;
        };
        using I22 = std::integral_constant<int, 22>; using I26 = std::integral_constant<int, 26>; using I82 = std::integral_constant<int, 82>;
        using O4 = std::integral_constant<int, 4>; using O8 = std::integral_constant<int, 8>;
        const bool o8 = h->pol.output_dim == 8;
        if(h->pol.input_dim == 22) rc = o8 ? go(I22{}, O8{}) : go(I22{}, O4{});
        else if(h->pol.input_dim == 26) rc = o8 ? go(I26{}, O8{}) : go(I26{}, O4{});
        else rc = o8 ? go(I82{}, O8{}) : go(I82{}, O4{});
        if(rc) return rc;
    return B200L2F_OK;
}
}  // namespace b200l2f
