"""raptor_b200 -- B200-native vectorised quadrotor rollout engine (l2f step/observe fused with the actor forward).

Host-side mirror of the reference's interface for this path:
  raptor_b200.VectorEnvironment   the l2f `vector::` surface (initialize_rng / initialize_environment / sample_initial_parameters /
                                  sample_initial_state / observe / step) + the fused `rollout`
  raptor_b200.l2f                 README-compatible module (l2f.vectorN.*, l2f.Device) -- R/README.md:40-105
  raptor_b200.foundation_policy   README-compatible module (Raptor().reset() / .evaluate_step()) -- R/README.md:19-24
Everything computes on CUDA through the C ABI in include/b200_l2f.h; importing works without a GPU, creating an environment does not.
"""
from ._lib import (DEVICE, FLAG_ACCURATE_MATH, GEMM_FP32_CUDA_CORES, GEMM_TCGEN05_3XTF32, HEAD_IDENTITY, HEAD_PPO_GAUSSIAN, HEAD_SQUASH_EVAL, HOST,  # noqa: F401
                   PARAMS_DIM, POLICY_MLP, POLICY_RAPTOR_GRU, SPEC_DEFAULT, SPEC_DEFAULT_DR, SPEC_RAPTOR, SPEC_RAPTOR_DR, SPEC_TEACHER, SPEC_TEACHER_DR)
from .engine import Checkpoint, EngineError, VectorEnvironment, parameters_from_json, parameters_to_json, raptor_policy_blob  # noqa: F401

__all__ = ["VectorEnvironment", "Checkpoint", "EngineError", "raptor_policy_blob", "parameters_to_json", "parameters_from_json"]
