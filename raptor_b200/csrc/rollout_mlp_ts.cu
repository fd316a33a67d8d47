// raptor_b200/csrc/rollout_mlp_ts.cu -- instantiations of k_rollout_mlp_ts (mlp_tc.cuh): MLP actors on tcgen05, and the host-side operand image.
#include "launch.h"
#include "mlp_tc.cuh"

namespace b200l2f {
namespace {
template <class Spec, int OUT, bool UNIFORM, bool AXIAL, bool NOISE = false>
int launch_rollout_mlp_ts(b200l2f_handle* h, RolloutArgs a){
    constexpr int IN = Spec::OBS_DIM;
    using SM = MlpTsSmem<IN, OUT>;
    auto kern = k_rollout_mlp_ts<Spec, OUT, UNIFORM, AXIAL, NOISE>;
    static bool configured[8] = {}; static int capacity[8] = {};
    int dev = h->cfg.device & 7;
    if(!configured[dev]){
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM::TOTAL_ROLLOUT));
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int sms = 0;
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
        capacity[dev] = 2 * sms;             // design point: ~92 KB smem and 256 TMEM columns per CTA -> 2 CTAs/SM
        configured[dev] = true;
    }
    int grid = 0, rc;
    if((rc = prepare_schedule(h, a, capacity[dev], &grid))) return rc;
    kern<<<grid, BLOCK, SM::TOTAL_ROLLOUT, h->stream>>>(a, h->d_mlp_tc_image);
    h->last_kernel = "k_rollout_mlp_ts";
    LAUNCH_CHECK();
    return B200L2F_OK;
}
}  // namespace

int launch_mlp_ts(b200l2f_handle* h, const RolloutArgs& a, bool uniform, bool axial, bool noise){
    const bool o8 = h->pol.output_dim == 8;
    if(noise){
        if(!uniform) return fail(h, B200L2F_ERR_UNSUPPORTED, "rollout: the tcgen05 noise variant needs uniform MDP constants");
        auto gon = [&](auto spec) -> int {
            using Spec = decltype(spec);
            if(axial) return o8 ? launch_rollout_mlp_ts<Spec, 8, true, true, true>(h, a) : launch_rollout_mlp_ts<Spec, 4, true, true, true>(h, a);
            return o8 ? launch_rollout_mlp_ts<Spec, 8, true, false, true>(h, a) : launch_rollout_mlp_ts<Spec, 4, true, false, true>(h, a);
        };
        return h->kind == KIND_RAPTOR ? gon(SpecRaptor{}) : gon(SpecTeacher{});
    }
            auto gots = [&](auto spec) -> int {
                using Spec = decltype(spec);
                if(!uniform) return o8 ? launch_rollout_mlp_ts<Spec, 8, false, false>(h, a) : launch_rollout_mlp_ts<Spec, 4, false, false>(h, a);
                if(axial) return o8 ? launch_rollout_mlp_ts<Spec, 8, true, true>(h, a) : launch_rollout_mlp_ts<Spec, 4, true, true>(h, a);
                return o8 ? launch_rollout_mlp_ts<Spec, 8, true, false>(h, a) : launch_rollout_mlp_ts<Spec, 4, true, false>(h, a);
            };
    return h->kind == KIND_RAPTOR ? gots(SpecRaptor{}) : gots(SpecTeacher{});
}

int build_mlp_tc_image(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob){
            auto build = [&](auto in_c, auto out_c) -> int {
                constexpr int IN = decltype(in_c)::value, OUT = decltype(out_c)::value;
                std::vector<float> img(MlpTcImage<IN, OUT>::SIZE);
                build_mlp_tc_image_host<IN, OUT>(img.data(), blob, desc->standardize != 0, desc->head == B200L2F_HEAD_PPO_GAUSSIAN);
                CU(cudaMalloc(&h->d_mlp_tc_image, MlpTcImage<IN, OUT>::BYTES));
                CU(cudaMemcpy(h->d_mlp_tc_image, img.data(), MlpTcImage<IN, OUT>::BYTES, cudaMemcpyHostToDevice));
                return (int)B200L2F_OK;
            };
            using I22 = std::integral_constant<int, 22>; using I26 = std::integral_constant<int, 26>;
            using O4 = std::integral_constant<int, 4>; using O8 = std::integral_constant<int, 8>;
            int brc;
            if(h->obs_dim == 82){   // DEFAULT spec: the PPO actor shape only (collection); squash-head rollouts stay on the CUDA-core kernels
                if(desc->output_dim != 4) return B200L2F_OK;
                brc = build(std::integral_constant<int, 82>{}, O4{});
            }
            else if(h->obs_dim == 22) brc = desc->output_dim == 8 ? build(I22{}, O8{}) : build(I22{}, O4{});
            else brc = desc->output_dim == 8 ? build(I26{}, O8{}) : build(I26{}, O4{});
            if(brc) return brc;
    return B200L2F_OK;
}
}  // namespace b200l2f
