"""ctypes binding of the C ABI declared in include/b200_l2f.h.  Loading fails loudly if the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200L2F_LIB") or os.path.join(HERE, "lib", "libb200l2f.so")   # B200L2F_LIB: tuning experiments only

c_int, c_i32, c_i64, c_u64, c_f = ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float
vp = ctypes.c_void_p

HOST, DEVICE = 0, 1
SPEC_DEFAULT, SPEC_DEFAULT_DR, SPEC_RAPTOR, SPEC_TEACHER, SPEC_RAPTOR_DR, SPEC_TEACHER_DR = range(6)
POLICY_RAPTOR_GRU, POLICY_MLP = 0, 1
HEAD_IDENTITY, HEAD_SQUASH_EVAL, HEAD_PPO_GAUSSIAN = 0, 1, 2
GEMM_FP32_CUDA_CORES, GEMM_TCGEN05_3XTF32 = 0, 1
FLAG_ACCURATE_MATH = 1
PARAMS_DIM = 145


class Config(ctypes.Structure):
    _fields_ = [("struct_size", c_i32), ("spec", c_i32), ("n_envs", c_i32), ("device", c_i32), ("first_env_id", c_i64),
                ("n_state_slots", c_i32), ("flags", c_i32), ("stream", vp)]


class PolicyDesc(ctypes.Structure):
    _fields_ = [("arch", c_i32), ("input_dim", c_i32), ("hidden_dim", c_i32), ("output_dim", c_i32), ("standardize", c_i32), ("head", c_i32),
                ("gru_sequence_length", c_i32), ("gemm", c_i32)]


class RolloutOut(ctypes.Structure):
    _fields_ = [("memspace", c_i32), ("state_stride", c_i32), ("states", vp), ("observations", vp), ("actions", vp), ("rewards", vp),
                ("terminated", vp), ("returns", vp), ("episode_length", vp)]


class ReplayBuffers(ctypes.Structure):
    _fields_ = [("memspace", c_i32), ("capacity", c_i32), ("data", vp), ("episode_start", vp), ("position", vp), ("full", vp), ("current_episode_start", vp)]


class Batch(ctypes.Structure):
    _fields_ = [("memspace", c_i32), ("batch_size", c_i32), ("observations_actions", vp), ("rewards", vp), ("terminated", vp), ("reset", vp), ("next_reset", vp),
                ("final_step_mask", vp), ("next_final_step_mask", vp), ("env_index", vp), ("sample_index", vp)]


class BatchParameters(ctypes.Structure):
    _fields_ = [("sequence_length", c_i32), ("include_first_step_in_targets", c_i32), ("always_sample_from_initial_state", c_i32), ("random_seq_length", c_i32),
                ("enable_nominal_sequence_length_probability", c_i32), ("nominal_sequence_length_probability", ctypes.c_float)]


class Status(ctypes.Structure):
    _fields_ = [("n_envs", c_i64), ("n_nonfinite", c_i64), ("n_terminated", c_i64), ("has_episodes", c_i32), ("reserved", c_i32),
                ("returns_mean", ctypes.c_double), ("returns_std", ctypes.c_double), ("episode_length_mean", ctypes.c_double), ("episode_length_std", ctypes.c_double),
                ("share_terminated", ctypes.c_double)]


class DaggerOut(ctypes.Structure):
    _fields_ = [("memspace", c_i32), ("reserved", c_i32), ("capacity_rows", c_i64), ("input_student", vp), ("output_target", vp), ("truncated", vp),
                ("reset", vp), ("episode_start", vp), ("returns", vp), ("episode_length", vp)]


# every symbol include/b200_l2f.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200l2f_create": (c_int, [ctypes.POINTER(Config), ctypes.POINTER(vp)]),
    "b200l2f_destroy": (c_int, [vp]),
    "b200l2f_last_error": (ctypes.c_char_p, [vp]),
    "b200l2f_synchronize": (c_int, [vp]),
    "b200l2f_stream": (vp, [vp]),
    "b200l2f_last_kernel": (ctypes.c_char_p, [vp]),
    "b200l2f_step_repeated": (c_int, [vp, c_int, vp, c_i32]),
    "b200l2f_allgather_trajectories": (c_int, [vp, vp, vp, vp, ctypes.c_size_t, ctypes.POINTER(c_i32)]),
    "b200l2f_last_status": (c_int, [vp, ctypes.POINTER(Status), vp, c_int]),
    "b200l2f_state_dim": (c_int, [vp]),
    "b200l2f_observation_dim": (c_int, [vp]),
    "b200l2f_action_history_length": (c_int, [vp]),
    "b200l2f_n_envs": (c_int, [vp]),
    "b200l2f_kernel_launches": (c_i64, [vp]),
    "b200l2f_initialize_rng": (c_int, [vp, c_u64, c_i32]),
    "b200l2f_get_rng": (c_int, [vp, vp, c_int]),
    "b200l2f_set_rng": (c_int, [vp, vp, c_int]),
    "b200l2f_initialize_environment": (c_int, [vp]),
    "b200l2f_get_environment_parameters": (c_int, [vp, vp]),
    "b200l2f_set_environment_parameters": (c_int, [vp, vp]),
    "b200l2f_initial_parameters": (c_int, [vp]),
    "b200l2f_sample_initial_parameters": (c_int, [vp]),
    "b200l2f_get_parameters": (c_int, [vp, vp, c_int]),
    "b200l2f_set_parameters": (c_int, [vp, vp, c_int]),
    "b200l2f_initial_state": (c_int, [vp, c_int]),
    "b200l2f_sample_initial_state": (c_int, [vp, c_int]),
    "b200l2f_get_state": (c_int, [vp, c_int, vp, c_int]),
    "b200l2f_set_state": (c_int, [vp, c_int, vp, c_int]),
    "b200l2f_set_parameters_async": (c_int, [vp, vp]),
    "b200l2f_set_state_async": (c_int, [vp, c_int, vp]),
    "b200l2f_get_state_async": (c_int, [vp, c_int, vp]),
    "b200l2f_copy_to_host_async": (c_int, [vp, vp, vp, ctypes.c_size_t]),
    "b200l2f_transfers_synchronize": (c_int, [vp, c_int]),
    "b200l2f_copy_state": (c_int, [vp, c_int, c_int]),
    "b200l2f_observe": (c_int, [vp, c_int, vp, c_int, c_int]),
    "b200l2f_step": (c_int, [vp, c_int, vp, c_int, vp, c_int]),
    "b200l2f_reward": (c_int, [vp, c_int, vp, c_int, vp, c_int]),
    "b200l2f_terminated": (c_int, [vp, c_int, vp, c_int]),
    "b200l2f_policy_load": (c_int, [vp, ctypes.POINTER(PolicyDesc), vp, ctypes.c_size_t]),
    "b200l2f_policy_reset": (c_int, [vp, vp, c_int]),
    "b200l2f_policy_evaluate_step": (c_int, [vp, vp, c_int, vp, c_int, c_int]),
    "b200l2f_policy_get_hidden": (c_int, [vp, vp, vp, c_int]),
    "b200l2f_policy_set_hidden": (c_int, [vp, vp, vp, c_int]),
    "b200l2f_rollout": (c_int, [vp, c_i32, c_i32, ctypes.POINTER(RolloutOut)]),
    "b200l2f_collect_reset": (c_int, [vp]),
    "b200l2f_collect": (c_int, [vp, c_i32, c_i32, vp, c_int]),
    "b200l2f_critic_load": (c_int, [vp, ctypes.POINTER(PolicyDesc), vp, ctypes.c_size_t]),
    "b200l2f_evaluate_values": (c_int, [vp, c_i32, vp, c_int]),
    "b200l2f_estimate_generalized_advantages": (c_int, [vp, c_i32, c_f, c_f, c_int, vp, c_int]),
    "b200l2f_values_and_advantages": (c_int, [vp, c_i32, c_f, c_f, c_int, vp, c_int]),
    "b200l2f_normalizer_update": (c_int, [vp, c_i32, vp, c_int, vp, vp, vp]),
    "b200l2f_parameters_to_json": (c_int, [vp, vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b200l2f_parameters_from_json": (c_int, [vp, ctypes.c_char_p, vp]),
    "b200l2f_state_to_json": (c_int, [vp, vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b200l2f_state_from_json": (c_int, [vp, ctypes.c_char_p, vp]),
    "b200l2f_checkpoint_parse": (c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "b200l2f_checkpoint_parse_h5": (c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(vp)]),
    "b200l2f_checkpoint_free": (c_int, [vp]),
    "b200l2f_checkpoint_tensor_count": (c_int, [vp]),
    "b200l2f_checkpoint_tensor": (c_int, [vp, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_i32), ctypes.POINTER(ctypes.POINTER(c_i64)), ctypes.POINTER(ctypes.POINTER(c_f))]),
    "b200l2f_checkpoint_string": (ctypes.c_char_p, [vp, ctypes.c_char_p]),
    "b200l2f_checkpoint_string_count": (c_int, [vp]),
    "b200l2f_checkpoint_string_at": (c_int, [vp, c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p)]),
    "b200l2f_checkpoint_policy": (c_int, [vp, ctypes.c_char_p, ctypes.POINTER(PolicyDesc), vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b200l2f_off_policy_steps": (c_int, [vp, c_i32, c_i32, c_i32, ctypes.POINTER(ReplayBuffers)]),
    "b200l2f_gather_batch": (c_int, [vp, ctypes.POINTER(ReplayBuffers), c_i32, c_i32, c_i32, vp, ctypes.POINTER(Batch)]),
    "b200l2f_gather_batch_sequential": (c_int, [vp, ctypes.POINTER(ReplayBuffers), ctypes.POINTER(BatchParameters), c_i32, c_i32, c_i32, vp, ctypes.POINTER(Batch)]),
    "b200l2f_runner_get_state": (c_int, [vp, vp, vp, vp, c_int]),
    "b200l2f_runner_set_state": (c_int, [vp, vp, vp, vp, c_int]),
    "b200l2f_teachers_load": (c_int, [vp, c_i32, c_i32, vp, vp, c_i32]),
    "b200l2f_dagger_gather": (c_int, [vp, c_i32, c_i32, ctypes.POINTER(DaggerOut), ctypes.POINTER(c_i64)]),
}

_lib = None


def load():
    """dlopen the engine; raises if the CUDA extension has not been built (python -m raptor_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("raptor_b200: %s is missing -- build it with `python -m raptor_b200.build` (CUDA-only engine, no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
