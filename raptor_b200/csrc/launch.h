// raptor_b200/csrc/launch.h -- non-template entry points of the translation units that hold the heavy kernel instantiations.
#pragma once
#include "handle.h"
#include "mlp.cuh"   // CollectArgs

namespace b200l2f {
// rollout_fp32.cu: k_rollout_raptor (actor on fp32 CUDA cores; carries the NOISE variant)
int launch_raptor_fp32(b200l2f_handle* h, const RolloutArgs& a, bool noise, bool fast, bool constw, bool rolled);
// rollout_tc.cu: k_rollout_raptor_tc (tcgen05, A operand in shared memory)
int launch_raptor_tc(b200l2f_handle* h, const RolloutArgs& a, bool fast, bool uniform, bool g1_tc);
// rollout_ts.cu: k_rollout_raptor_ts (tcgen05, A operand + hidden state in TMEM; the default hot path)
int launch_raptor_ts(b200l2f_handle* h, const RolloutArgs& a, bool uniform, bool axial, bool noise);   // noise requires uniform
// rollout_x2.cu: k_rollout_raptor_x2 (two environments per thread on the packed fp32 pipe; RAPTOR spec, default math, uniform MDP constants, axial vehicles, no noise)
int launch_raptor_x2(b200l2f_handle* h, const RolloutArgs& a);
// rollout_mlp.cu: k_rollout_mlp, k_mlp_step (MLP actors on CUDA cores)
int launch_mlp_fp32(b200l2f_handle* h, const RolloutArgs& a);
int launch_mlp_step(b200l2f_handle* h, const float* d_obs, int ld, float* d_act);
// rollout_mlp_ts.cu: k_rollout_mlp_ts (MLP actors on tcgen05) + the operand image
int launch_mlp_ts(b200l2f_handle* h, const RolloutArgs& a, bool uniform, bool axial, bool noise);      // noise requires uniform
int build_mlp_tc_image(b200l2f_handle* h, const b200l2f_policy_desc* desc, const float* blob);
// collect.cu / collect_ts.cu: k_collect, k_collect_ts
int launch_collect_fp32(b200l2f_handle* h, const CollectArgs& a);
int launch_collect_ts(b200l2f_handle* h, const CollectArgs& a, bool follow, bool row_axial);
// collect_ts_default.cu: k_collect_ts for the DEFAULT spec (H = 16 action ring, OBS 82: the first layer is a K = 88 operand, one CTA per SM)
int launch_collect_ts_default(b200l2f_handle* h, const CollectArgs& a, bool follow, bool row_axial);
}  // namespace b200l2f
