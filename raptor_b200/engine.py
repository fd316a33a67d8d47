"""Host-side mirror of the reference's vector:: environment interface on top of the C ABI (include/b200_l2f.h).

Arguments may be numpy float32 arrays (host memory: copies are staged through pinned memory inside the call) or torch CUDA tensors
(device memory: zero-copy, work is enqueued on the engine's stream; tensors produced on torch's current stream are ordered before it automatically,
results must be awaited with synchronize() -- or create the environment on torch's stream: VectorEnvironment(..., stream=torch.cuda.current_stream().cuda_stream)).  Names, argument meaning and error behaviour follow the reference:
  rl_tools::init / initial_parameters / sample_initial_parameters / initial_state / sample_initial_state / observe / step / reward /
  terminated   (rl_tools/rl/environments/l2f/operations_generic.h:43-176)
  rl_tools::reset / evaluate_step  (rl_tools/nn_models/sequential/operations_generic.h:63-66,321-325)
  rl_tools::evaluate               (rl_tools/rl/utils/evaluation/operations_generic.h:93-214)  -> VectorEnvironment.rollout
"""
import ctypes
import os

import numpy as np

from . import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))


class EngineError(RuntimeError):
    pass


def raptor_policy_blob():
    """the Raptor checkpoint's 2084 weights in the engine's blob layout (extracted from the reference checkpoint by tests/golden/generate.py)"""
    return np.fromfile(os.path.join(HERE, "data", "raptor_policy_2084.f32"), dtype=np.float32)


class Checkpoint:
    """an rl-tools checkpoint -- the code export `checkpoint.h` (rl/loop/steps/checkpoint/operations_cpu.h:56-118) or its HDF5 twin `checkpoint.h5`
    (:119-160; told apart by the HDF5 signature) -- read by the engine's native readers (csrc/checkpoint_io.cu, csrc/h5_io.cu); no GPU needed.
    `tensors` maps namespace paths to float32 arrays, `strings` the meta strings / HDF5 string attributes; `policy()` returns (PolicyDesc, blob) for
    VectorEnvironment.load_policy(**Checkpoint.policy_kwargs())."""

    def __init__(self, text=None, path=None):
        lib = L.load()
        if text is None:
            with open(path, "rb") as f:
                text = f.read()
        if isinstance(text, str):
            text = text.encode()
        h = ctypes.c_void_p()
        if lib.b200l2f_checkpoint_parse(text, len(text), ctypes.byref(h)) != 0:
            raise EngineError(lib.b200l2f_last_error(None).decode())
        try:
            self.tensors = {}
            for i in range(lib.b200l2f_checkpoint_tensor_count(h)):
                name, rank = ctypes.c_char_p(), ctypes.c_int32()
                dims, data = ctypes.POINTER(ctypes.c_int64)(), ctypes.POINTER(ctypes.c_float)()
                lib.b200l2f_checkpoint_tensor(h, i, ctypes.byref(name), ctypes.byref(rank), ctypes.byref(dims), ctypes.byref(data))
                shape = tuple(dims[k] for k in range(rank.value))
                count = int(np.prod(shape, dtype=np.int64))
                self.tensors[name.value.decode()] = (np.ctypeslib.as_array(data, shape=(count,)).copy() if count else np.zeros(0, np.float32)).reshape(shape)
            self.strings = {}
            for i in range(lib.b200l2f_checkpoint_string_count(h)):
                k, v = ctypes.c_char_p(), ctypes.c_char_p()
                lib.b200l2f_checkpoint_string_at(h, i, ctypes.byref(k), ctypes.byref(v))
                self.strings[k.value.decode()] = v.value.decode(errors="replace")
            self.name, self.commit_hash = self.string("rl_tools::checkpoint::meta::name"), self.string("rl_tools::checkpoint::meta::commit_hash")
            self._policy = {}
            self._errors = {}
            for root in ("rl_tools::checkpoint::actor",):
                desc, n = L.PolicyDesc(), ctypes.c_size_t(0)
                if lib.b200l2f_checkpoint_policy(h, root.encode(), ctypes.byref(desc), None, 0, ctypes.byref(n)) == 0:
                    blob = np.zeros(n.value, np.float32)
                    lib.b200l2f_checkpoint_policy(h, root.encode(), ctypes.byref(desc), blob.ctypes.data, blob.size, ctypes.byref(n))
                    self._policy[root] = (desc, blob)
                else:
                    self._errors[root] = lib.b200l2f_last_error(None).decode()
        finally:
            lib.b200l2f_checkpoint_free(h)

    def string(self, path):
        """`char name[] = "..."` of the code export / string attribute of the .h5 ("<namespace path>::<name>"), or None"""
        return self.strings.get(path)

    def policy(self, root="rl_tools::checkpoint::actor"):
        if root not in self._policy:
            raise EngineError(self._errors.get(root, "checkpoint: no actor under " + root))
        return self._policy[root]

    def policy_kwargs(self, root="rl_tools::checkpoint::actor"):
        d, blob = self.policy(root)
        return dict(blob=blob, arch=d.arch, input_dim=d.input_dim, hidden_dim=d.hidden_dim, output_dim=d.output_dim, standardize=d.standardize,
                    head=d.head, gru_sequence_length=d.gru_sequence_length, gemm=d.gemm)

    @property
    def example(self):
        """the known-answer pair stored with the actor (input, output) or None"""
        i, o = self.tensors.get("rl_tools::checkpoint::example::input"), self.tensors.get("rl_tools::checkpoint::example::output")
        return None if i is None or o is None else (i, o)


def parameters_to_json(row):
    """rl_tools::json(device, env, parameters) (rl/environments/l2f/operations_cpu.h:139-411) of one flat parameter row [145]; no GPU needed"""
    lib = L.load()
    row = np.ascontiguousarray(row, np.float32)
    if row.shape != (L.PARAMS_DIM,):
        raise ValueError("parameters_to_json: expected a row of %d floats" % L.PARAMS_DIM)
    n = ctypes.c_size_t(0)
    lib.b200l2f_parameters_to_json(None, row.ctypes.data, None, 0, ctypes.byref(n))
    buf = ctypes.create_string_buffer(n.value + 1)
    rc = lib.b200l2f_parameters_to_json(None, row.ctypes.data, buf, n.value + 1, ctypes.byref(n))
    if rc != 0:
        raise EngineError("b200l2f_parameters_to_json failed (%d): %s" % (rc, lib.b200l2f_last_error(None).decode()))
    return buf.value.decode()


def parameters_from_json(text, row=None):
    """rl_tools::from_json (operations_cpu.h:565-728): returns the flat row [145]; `row` supplies the values of keys the format does not carry
    (none today) and is not modified; no GPU needed"""
    lib = L.load()
    out = np.zeros(L.PARAMS_DIM, np.float32) if row is None else np.array(row, np.float32, copy=True)
    rc = lib.b200l2f_parameters_from_json(None, text.encode(), out.ctypes.data)
    if rc != 0:
        raise EngineError("b200l2f_parameters_from_json failed (%d): %s" % (rc, lib.b200l2f_last_error(None).decode()))
    return out


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _arg(x, dtype, shape=None, name="array", engine_stream=None):
    """-> (pointer, memspace, keepalive).  engine_stream: the cudaStream_t the engine enqueues on; a torch tensor produced on another stream
    (torch's current one) is ordered before the engine's work by synchronising that stream first (no-op when the streams are the same)."""
    if x is None:
        return None, L.HOST, None
    if _is_torch(x):
        import torch
        if engine_stream is not None and x.is_cuda:
            cur = torch.cuda.current_stream(x.device)
            if (cur.cuda_stream or 0) != (engine_stream or 0):
                cur.synchronize()
        want = {np.float32: torch.float32, np.uint64: torch.int64, np.int32: torch.int32, np.uint8: torch.uint8}[dtype]
        if x.dtype != want and not (dtype == np.uint64 and x.dtype == getattr(torch, "uint64", None)):
            raise TypeError("%s: expected torch dtype %s, got %s" % (name, want, x.dtype))
        if not x.is_cuda or not x.is_contiguous():
            raise ValueError("%s: torch tensors must be contiguous CUDA tensors (use numpy arrays for host data)" % name)
        if shape is not None and tuple(x.shape) != tuple(shape):
            raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(x.shape)))
        return ctypes.c_void_p(x.data_ptr()), L.DEVICE, x
    a = x
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError("%s: expected a C-contiguous numpy array of dtype %s" % (name, np.dtype(dtype)))
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(a.shape)))
    return ctypes.c_void_p(a.ctypes.data), L.HOST, a


class VectorEnvironment:
    """N quadrotor environments on one GPU (one handle = one shard; shards never communicate)."""

    def __init__(self, n_envs, spec=L.SPEC_DEFAULT, device=0, first_env_id=0, n_state_slots=2, flags=0, stream=None):
        self._lib = L.load()
        cfg = L.Config(ctypes.sizeof(L.Config), spec, n_envs, device, first_env_id, n_state_slots, flags, stream)
        h = ctypes.c_void_p()
        rc = self._lib.b200l2f_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise EngineError("b200l2f_create failed (%d): %s" % (rc, self._lib.b200l2f_last_error(None).decode()))
        self._h = h
        self.spec = spec
        self.device = device
        self.N_ENVIRONMENTS = n_envs
        self.OBSERVATION_DIM = self._lib.b200l2f_observation_dim(h)
        self.STATE_DIM = self._lib.b200l2f_state_dim(h)
        self.ACTION_DIM = 4
        self.ACTION_HISTORY_LENGTH = self._lib.b200l2f_action_history_length(h)
        self.policy = None
        self._engine_stream = self._lib.b200l2f_stream(h)

    # ---- plumbing
    def _arg(self, x, dtype, shape=None, name="array"):
        return _arg(x, dtype, shape, name, engine_stream=self._engine_stream)

    def _check(self, rc):
        if rc != 0:
            raise EngineError("b200l2f error %d: %s" % (rc, self._lib.b200l2f_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200l2f_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self._lib.b200l2f_synchronize(self._h))

    @property
    def stream(self):
        return self._lib.b200l2f_stream(self._h)

    def last_status(self, flags=False):
        """status of the last rollout / collect: dict(n_nonfinite, and for rollouts the rl::utils::evaluation::Result aggregates returns_mean / returns_std /
        episode_length_mean / episode_length_std / n_terminated / share_terminated); flags=True adds the per-environment non-finite flags"""
        st = L.Status()
        f = np.zeros(self.N_ENVIRONMENTS, np.uint8) if flags else None
        self._check(self._lib.b200l2f_last_status(self._h, ctypes.byref(st), f.ctypes.data if flags else None, L.HOST))
        out = {k: getattr(st, k) for k, _ in L.Status._fields_ if k != "reserved"}
        if flags:
            out["nonfinite_flags"] = f
        return out

    def last_kernel(self):
        return self._lib.b200l2f_last_kernel(self._h).decode()

    @property
    def kernel_launches(self):
        return int(self._lib.b200l2f_kernel_launches(self._h))

    # ---- RNG / environment / parameters
    def initialize_rng(self, seed=0, warmup=0):
        self._check(self._lib.b200l2f_initialize_rng(self._h, seed, warmup))

    def get_rng(self):
        out = np.zeros(self.N_ENVIRONMENTS, np.uint64)
        self._check(self._lib.b200l2f_get_rng(self._h, out.ctypes.data, L.HOST))
        return out

    def set_rng(self, states):
        p, ms, _ = self._arg(states, np.uint64, (self.N_ENVIRONMENTS,), "rng")
        self._check(self._lib.b200l2f_set_rng(self._h, p, ms))

    def initialize_environment(self):
        self._check(self._lib.b200l2f_initialize_environment(self._h))

    def get_environment_parameters(self):
        row = np.zeros(L.PARAMS_DIM, np.float32)
        self._check(self._lib.b200l2f_get_environment_parameters(self._h, row.ctypes.data))
        return row

    def set_environment_parameters(self, row):
        row = np.ascontiguousarray(row, np.float32)
        assert row.shape == (L.PARAMS_DIM,)
        self._check(self._lib.b200l2f_set_environment_parameters(self._h, row.ctypes.data))

    def initial_parameters(self):
        self._check(self._lib.b200l2f_initial_parameters(self._h))

    def sample_initial_parameters(self):
        self._check(self._lib.b200l2f_sample_initial_parameters(self._h))

    def get_parameters(self, out=None):
        out = np.zeros((self.N_ENVIRONMENTS, L.PARAMS_DIM), np.float32) if out is None else out
        p, ms, _ = self._arg(out, np.float32, (self.N_ENVIRONMENTS, L.PARAMS_DIM), "parameters")
        self._check(self._lib.b200l2f_get_parameters(self._h, p, ms))
        return out

    def set_parameters(self, rows):
        p, ms, _ = self._arg(rows, np.float32, (self.N_ENVIRONMENTS, L.PARAMS_DIM), "parameters")
        self._check(self._lib.b200l2f_set_parameters(self._h, p, ms))

    # ---- state
    def initial_state(self, slot=0):
        self._check(self._lib.b200l2f_initial_state(self._h, slot))

    def sample_initial_state(self, slot=0):
        self._check(self._lib.b200l2f_sample_initial_state(self._h, slot))

    def get_state(self, slot=0, out=None):
        out = np.zeros((self.N_ENVIRONMENTS, self.STATE_DIM), np.float32) if out is None else out
        p, ms, _ = self._arg(out, np.float32, (self.N_ENVIRONMENTS, self.STATE_DIM), "state")
        self._check(self._lib.b200l2f_get_state(self._h, slot, p, ms))
        return out

    def set_state(self, rows, slot=0):
        p, ms, _ = self._arg(rows, np.float32, (self.N_ENVIRONMENTS, self.STATE_DIM), "state")
        self._check(self._lib.b200l2f_set_state(self._h, slot, p, ms))

    # ---- asynchronous transfers (page-locked host arrays, e.g. numpy views of torch pinned tensors): see include/b200_l2f.h
    def _pinned_arg(self, a, shape, name):
        a = np.asarray(a)
        if a.dtype != np.float32 or tuple(a.shape) != tuple(shape) or not a.flags["C_CONTIGUOUS"]:
            raise ValueError("%s: expected a C-contiguous float32 array of shape %s" % (name, tuple(shape)))
        return a.ctypes.data

    def set_parameters_async(self, rows):
        self._check(self._lib.b200l2f_set_parameters_async(self._h, self._pinned_arg(rows, (self.N_ENVIRONMENTS, L.PARAMS_DIM), "parameters")))

    def set_state_async(self, rows, slot=0):
        self._check(self._lib.b200l2f_set_state_async(self._h, slot, self._pinned_arg(rows, (self.N_ENVIRONMENTS, self.STATE_DIM), "state")))

    def get_state_async(self, out, slot=0):
        self._check(self._lib.b200l2f_get_state_async(self._h, slot, self._pinned_arg(out, (self.N_ENVIRONMENTS, self.STATE_DIM), "state")))
        return out

    def copy_to_host_async(self, dst, src_device):
        """dst: page-locked numpy array; src_device: torch CUDA tensor the engine's stream produced (same byte size)"""
        dst = np.asarray(dst)
        nbytes = src_device.numel() * src_device.element_size()
        if dst.nbytes != nbytes or not dst.flags["C_CONTIGUOUS"]:
            raise ValueError("copy_to_host_async: destination must be C-contiguous with %d bytes" % nbytes)
        self._check(self._lib.b200l2f_copy_to_host_async(self._h, dst.ctypes.data, src_device.data_ptr(), nbytes))
        return dst

    def transfers_synchronize(self, uploads=True, downloads=True):
        self._check(self._lib.b200l2f_transfers_synchronize(self._h, (1 if uploads else 0) | (2 if downloads else 0)))

    def copy_state(self, dst_slot, src_slot):
        self._check(self._lib.b200l2f_copy_state(self._h, dst_slot, src_slot))

    # ---- observe / step / reward / terminated
    def observe(self, observation=None, slot=0):
        if observation is None:
            observation = np.zeros((self.N_ENVIRONMENTS, self.OBSERVATION_DIM), np.float32)
        if observation.shape[0] != self.N_ENVIRONMENTS or observation.shape[1] < self.OBSERVATION_DIM:
            raise ValueError("observe: observation must be [N_ENVIRONMENTS, >= OBSERVATION_DIM]")
        p, ms, _ = self._arg(observation, np.float32, None, "observation")
        self._check(self._lib.b200l2f_observe(self._h, slot, p, observation.shape[1], ms))
        return observation

    def step(self, action, slot=0, next_slot=1, dts=None):
        p, ms, _ = self._arg(action, np.float32, (self.N_ENVIRONMENTS, 4), "action")
        if dts is None and ms == L.HOST:
            dts = np.zeros(self.N_ENVIRONMENTS, np.float32)
        pd, msd, _ = self._arg(dts, np.float32, (self.N_ENVIRONMENTS,), "dts")
        if dts is not None and msd != ms:
            raise ValueError("step: action and dts must live in the same memory space")
        self._check(self._lib.b200l2f_step(self._h, slot, p, next_slot, pd, ms))
        return dts

    def step_repeated(self, action4, n_steps, slot=0):
        """n_steps x step under one held action for every environment, in place on `slot` (the loop of the reference's GPU benchmark)"""
        a = np.ascontiguousarray(action4, np.float32).reshape(4)
        self._check(self._lib.b200l2f_step_repeated(self._h, slot, a.ctypes.data, int(n_steps)))

    def reward(self, action, slot=0, next_slot=1, out=None):
        p, ms, _ = self._arg(action, np.float32, (self.N_ENVIRONMENTS, 4), "action")
        out = np.zeros(self.N_ENVIRONMENTS, np.float32) if out is None else out
        po, mso, _ = self._arg(out, np.float32, (self.N_ENVIRONMENTS,), "rewards")
        if mso != ms:
            raise ValueError("reward: action and out must live in the same memory space")
        self._check(self._lib.b200l2f_reward(self._h, slot, p, next_slot, po, ms))
        return out

    def terminated(self, slot=0, out=None):
        out = np.zeros(self.N_ENVIRONMENTS, np.uint8) if out is None else out
        p, ms, _ = self._arg(out, np.uint8, (self.N_ENVIRONMENTS,), "flags")
        self._check(self._lib.b200l2f_terminated(self._h, slot, p, ms))
        return out

    # ---- actor
    def load_policy(self, blob=None, arch=L.POLICY_RAPTOR_GRU, input_dim=22, hidden_dim=16, output_dim=4, standardize=0, head=L.HEAD_IDENTITY,
                    gru_sequence_length=500, gemm=L.GEMM_TCGEN05_3XTF32):
        blob = raptor_policy_blob() if blob is None else np.ascontiguousarray(blob, np.float32)
        desc = L.PolicyDesc(arch, input_dim, hidden_dim, output_dim, standardize, head, gru_sequence_length, gemm)
        self._check(self._lib.b200l2f_policy_load(self._h, ctypes.byref(desc), blob.ctypes.data, blob.size))
        self.policy = desc

    def policy_reset(self, mask=None):
        p, ms, _ = self._arg(mask, np.uint8, (self.N_ENVIRONMENTS,), "mask")
        self._check(self._lib.b200l2f_policy_reset(self._h, p, ms))

    def policy_evaluate_step(self, observation, action=None, no_auto_reset=False):
        po, ms, _ = self._arg(observation, np.float32, None, "observation")
        if observation.shape[0] != self.N_ENVIRONMENTS or observation.shape[1] < self.policy.input_dim:
            raise ValueError("evaluate_step: observation must be [N_ENVIRONMENTS, >= input_dim]")
        if action is None:
            if ms == L.HOST:
                action = np.zeros((self.N_ENVIRONMENTS, 4), np.float32)
            else:
                import torch
                action = torch.empty((self.N_ENVIRONMENTS, 4), dtype=torch.float32, device=observation.device)
        pa, msa, _ = self._arg(action, np.float32, (self.N_ENVIRONMENTS, 4), "action")
        if msa != ms:
            raise ValueError("evaluate_step: observation and action must live in the same memory space")
        self._check(self._lib.b200l2f_policy_evaluate_step(self._h, po, observation.shape[1], pa, int(no_auto_reset), ms))
        return action

    # ---- off-policy runner (rl::components::off_policy_runner step): SAC-teacher collection into per-environment replay rings
    def new_replay_buffers(self, capacity, device=False):
        """empty rings as after rl_tools::init(device, runner): numpy arrays, or torch CUDA tensors (updated in place, no staging) if device"""
        n, D = self.N_ENVIRONMENTS, 2 * self.OBSERVATION_DIM + 7
        shapes = dict(data=((n, capacity, D), np.float32), episode_start=((n, capacity), np.int32), position=((n,), np.int32), full=((n,), np.uint8),
                      current_episode_start=((n,), np.int32))
        if not device:
            return {k: np.zeros(sh, dt) for k, (sh, dt) in shapes.items()}
        import torch
        tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8}
        return {k: torch.zeros(sh, dtype=tdt[dt], device="cuda:%d" % self.device) for k, (sh, dt) in shapes.items()}

    def off_policy_steps(self, n_steps, episode_step_limit, replay, sample_parameters=True):
        n, D = self.N_ENVIRONMENTS, 2 * self.OBSERVATION_DIM + 7
        capacity = replay["data"].shape[1]
        pd, ms, _ = self._arg(replay["data"], np.float32, (n, capacity, D), "replay.data")
        ptrs = [pd]
        for k, dt, sh in (("episode_start", np.int32, (n, capacity)), ("position", np.int32, (n,)), ("full", np.uint8, (n,)), ("current_episode_start", np.int32, (n,))):
            p, m, _ = self._arg(replay[k], dt, sh, "replay." + k)
            if m != ms:
                raise ValueError("off_policy_steps: all replay buffers must live in the same memory space")
            ptrs.append(p)
        rb = L.ReplayBuffers(ms, capacity, *ptrs)
        self._check(self._lib.b200l2f_off_policy_steps(self._h, n_steps, episode_step_limit, int(sample_parameters), ctypes.byref(rb)))
        return replay

    def gather_batch(self, replay, rng_states, max_episode_length=500, env_begin=0, env_count=None, out=None, sequence_length=None, include_first_step_in_targets=None,
                     always_sample_from_initial_state=None, random_seq_length=None, enable_nominal_sequence_length_probability=True, nominal_sequence_length_probability=0.5):
        """batch (rl_tools::gather_batch) from the replay rings: rng_states [B] uint64 (numpy) / int64 (torch CUDA), advanced in place.
        sequence_length None: the SEQUENCE_LENGTH-1 MLP SAC configuration; returns dict(observations_actions [2, B, OBS+4], rewards [B], terminated, reset, next_reset [2, B],
        final_step_mask, next_final_step_mask [2, B], env_index, sample_index).
        sequence_length L: the general walk (recurrent SAC); the flags default to the reference's for the given length (off_policy_runner.h:78-85: all three = L > 1);
        tensors in the SequentialBatch shapes: observations_actions [L+1, B, OBS+4], rewards / terminated / reset / final_step_mask [L, B], next_* [L+1, B]."""
        n, obs = self.N_ENVIRONMENTS, self.OBSERVATION_DIM
        capacity = replay["data"].shape[1]
        pd, ms, _ = self._arg(replay["data"], np.float32, (n, capacity, 2 * obs + 7), "replay.data")
        pp, m1, _ = self._arg(replay["position"], np.int32, (n,), "replay.position")
        pf, m2, _ = self._arg(replay["full"], np.uint8, (n,), "replay.full")
        B = int(rng_states.shape[0])
        pr, m3, _ = self._arg(rng_states, np.uint64, (B,), "rng_states")
        if not (ms == m1 == m2 == m3):
            raise ValueError("gather_batch: rings and rng_states must live in the same memory space")
        sequential = sequence_length is not None
        Lq = int(sequence_length) if sequential else 1
        cur = (Lq, B) if sequential else (B,)
        shapes = dict(observations_actions=((Lq + 1, B, obs + 4), np.float32), rewards=(cur, np.float32), terminated=(cur, np.uint8), reset=(cur, np.uint8),
                      next_reset=((Lq + 1, B), np.uint8), final_step_mask=(cur, np.uint8), next_final_step_mask=((Lq + 1, B), np.uint8), env_index=((B,), np.int32),
                      sample_index=((B,), np.int32))
        if out is None:
            if ms == L.HOST:
                out = {k: np.zeros(sh, dt) for k, (sh, dt) in shapes.items()}
            else:
                import torch
                tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8}
                out = {k: torch.zeros(sh, dtype=tdt[dt], device=rng_states.device) for k, (sh, dt) in shapes.items()}
        ptrs = []
        for k, (sh, dt) in shapes.items():
            p, m, _ = self._arg(out.get(k), dt, sh, "batch." + k)
            if out.get(k) is not None and m != ms:
                raise ValueError("gather_batch: batch buffers must live in the memory space of the rings")
            ptrs.append(p)
        batch = L.Batch(ms, B, *ptrs)
        count = n if env_count is None else env_count
        if not sequential:
            rb = L.ReplayBuffers(ms, capacity, pd, None, pp, pf, None)
            self._check(self._lib.b200l2f_gather_batch(self._h, ctypes.byref(rb), max_episode_length, env_begin, count, pr, ctypes.byref(batch)))
            return out
        pe, m4, _ = self._arg(replay["episode_start"], np.int32, (n, capacity), "replay.episode_start")
        if m4 != ms:
            raise ValueError("gather_batch: rings and episode_start must live in the same memory space")
        dflt = Lq > 1
        flag = lambda v: int(dflt if v is None else bool(v))  # noqa: E731
        bp = L.BatchParameters(Lq, flag(include_first_step_in_targets), flag(always_sample_from_initial_state), flag(random_seq_length),
                               int(bool(enable_nominal_sequence_length_probability)), float(nominal_sequence_length_probability))
        rb = L.ReplayBuffers(ms, capacity, pd, pe, pp, pf, None)
        self._check(self._lib.b200l2f_gather_batch_sequential(self._h, ctypes.byref(rb), ctypes.byref(bp), max_episode_length, env_begin, count, pr, ctypes.byref(batch)))
        return out

    def get_runner_state(self):
        """(episode_step [N] int32, episode_return [N] float32, truncated [N] uint8) of the on-/off-policy runner bookkeeping"""
        n = self.N_ENVIRONMENTS
        s, r, t = np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
        self._check(self._lib.b200l2f_runner_get_state(self._h, s.ctypes.data, r.ctypes.data, t.ctypes.data, L.HOST))
        return s, r, t

    def set_runner_state(self, episode_step=None, episode_return=None, truncated=None):
        n = self.N_ENVIRONMENTS
        ps, _, k1 = self._arg(None if episode_step is None else np.ascontiguousarray(episode_step, np.int32), np.int32, (n,), "episode_step")
        pr, _, k2 = self._arg(None if episode_return is None else np.ascontiguousarray(episode_return, np.float32), np.float32, (n,), "episode_return")
        pt, _, k3 = self._arg(None if truncated is None else np.ascontiguousarray(truncated, np.uint8), np.uint8, (n,), "truncated")
        self._check(self._lib.b200l2f_runner_set_state(self._h, ps, pr, pt, L.HOST))

    # ---- PPO collection (rl_tools::collect): on-device auto-reset + trajectory write-back
    def collect_reset(self):
        self._check(self._lib.b200l2f_collect_reset(self._h))

    def collect(self, n_steps, episode_step_limit, dataset=None):
        """dataset: [(T+1)*N, OBS+15] rows = step*N + env, columns obs | actions_mean[4] | actions[4] | log_prob | reward | terminated |
        truncated | value | advantage | target_value (the last three are the learner's); numpy (host) or torch CUDA tensor"""
        D = self.OBSERVATION_DIM + 15
        shape = ((n_steps + 1) * self.N_ENVIRONMENTS, D)
        if dataset is None:
            dataset = np.zeros(shape, np.float32)
        p, ms, _ = self._arg(dataset, np.float32, shape, "dataset")
        self._check(self._lib.b200l2f_collect(self._h, n_steps, episode_step_limit, p, ms))
        return dataset

    # ---- PPO learner feed (what the reference's loop step does between collect and train, rl/algorithms/ppo/loop/core/operations_generic.h:104-117)
    def load_critic(self, blob, standardize=0, gemm=L.GEMM_TCGEN05_3XTF32):
        """value network [standardize ->] Dense(OBS,64,ReLU) -> Dense(64,64,ReLU) -> Dense(64,1), blob in the MLP order of include/b200_l2f.h"""
        blob = np.ascontiguousarray(blob, np.float32)
        desc = L.PolicyDesc(L.POLICY_MLP, self.OBSERVATION_DIM, 64, 1, standardize, L.HEAD_IDENTITY, 0, gemm)
        self._check(self._lib.b200l2f_critic_load(self._h, ctypes.byref(desc), blob.ctypes.data, blob.size))

    def _dataset_arg(self, dataset, n_steps):
        return self._arg(dataset, np.float32, ((n_steps + 1) * self.N_ENVIRONMENTS, self.OBSERVATION_DIM + 15), "dataset")

    def evaluate_values(self, dataset, n_steps):
        """critic over all (T+1) N observation rows -> the all_values column, in place"""
        p, ms, _ = self._dataset_arg(dataset, n_steps)
        self._check(self._lib.b200l2f_evaluate_values(self._h, n_steps, p, ms))
        return dataset

    def estimate_generalized_advantages(self, dataset, n_steps, gamma=0.99, lam=0.95, ignore_termination=False):
        """rl_tools::estimate_generalized_advantages (rl/algorithms/ppo/operations_generic.h:54-89) on the value column, in place"""
        p, ms, _ = self._dataset_arg(dataset, n_steps)
        self._check(self._lib.b200l2f_estimate_generalized_advantages(self._h, n_steps, gamma, lam, int(ignore_termination), p, ms))
        return dataset

    def values_and_advantages(self, dataset, n_steps, gamma=0.99, lam=0.95, ignore_termination=False):
        """evaluate_values + estimate_generalized_advantages in one backward pass over time (the dataset is read once)"""
        p, ms, _ = self._dataset_arg(dataset, n_steps)
        self._check(self._lib.b200l2f_values_and_advantages(self._h, n_steps, gamma, lam, int(ignore_termination), p, ms))
        return dataset

    def normalizer_update(self, dataset, n_steps, mean, std, age):
        """rl::components::running_normalizer update with the dataset's observation block; mean / std: host float32 [OBS], updated in place;
        returns the new age"""
        p, ms, _ = self._dataset_arg(dataset, n_steps)
        pm, _, _ = self._arg(mean, np.float32, (self.OBSERVATION_DIM,), "mean")
        ps, _, _ = self._arg(std, np.float32, (self.OBSERVATION_DIM,), "std")
        a = ctypes.c_int32(age)
        self._check(self._lib.b200l2f_normalizer_update(self._h, n_steps, p, ms, pm, ps, ctypes.byref(a)))
        return int(a.value)

    # ---- JSON wire format of the reference (rl/environments/l2f/operations_cpu.h)
    def state_to_json(self, state_row):
        row = np.ascontiguousarray(state_row, np.float32)
        n = ctypes.c_size_t(0)
        self._lib.b200l2f_state_to_json(self._h, row.ctypes.data, None, 0, ctypes.byref(n))
        buf = ctypes.create_string_buffer(n.value + 1)
        self._check(self._lib.b200l2f_state_to_json(self._h, row.ctypes.data, buf, n.value + 1, ctypes.byref(n)))
        return buf.value.decode()

    def state_from_json(self, text, state_row=None):
        out = np.zeros(self.STATE_DIM, np.float32) if state_row is None else np.array(state_row, np.float32, copy=True)
        self._check(self._lib.b200l2f_state_from_json(self._h, text.encode(), out.ctypes.data))
        return out

    # ---- foundation-policy DAgger data path (src/foundation_policy/post_training/helper.h: gather_epoch = sample_trajectories + add_to_dataset)
    def load_teachers(self, blobs, position_offsets=None, episodes_per_teacher=10, gemm=L.GEMM_TCGEN05_3XTF32):
        """blobs [n_teachers, 6408] (SAC actor MLP 26-64-64-8), position_offsets [n_teachers, 3] or None; environment e is an episode of teacher
        e // episodes_per_teacher"""
        blobs = np.ascontiguousarray(blobs, np.float32)
        n_teachers = blobs.shape[0]
        off = None if position_offsets is None else np.ascontiguousarray(position_offsets, np.float32)
        if blobs.ndim != 2 or blobs.shape[1] != 6408 or (off is not None and off.shape != (n_teachers, 3)):
            raise ValueError("load_teachers: blobs [n_teachers, 6408], position_offsets [n_teachers, 3]")
        self._check(self._lib.b200l2f_teachers_load(self._h, n_teachers, episodes_per_teacher, blobs.ctypes.data, None if off is None else off.ctypes.data, gemm))

    def dagger_gather(self, n_steps, out=None, no_auto_reset=False):
        """student rollout + dataset rows of all teachers; returns dict(rows, input_student, output_target, truncated, reset, episode_start,
        returns, episode_length).  out: dict of preallocated buffers (numpy, or torch CUDA tensors for a device-resident dataset)"""
        n, cap = self.N_ENVIRONMENTS, self.N_ENVIRONMENTS * n_steps
        if out is None:
            out = dict(input_student=np.zeros((cap, 22), np.float32), output_target=np.zeros((cap, 4), np.float32), truncated=np.zeros(cap, np.uint8),
                       reset=np.zeros(cap, np.uint8), episode_start=np.zeros(n, np.int32), returns=np.zeros(n, np.float32), episode_length=np.zeros(n, np.int32))
        dt = dict(input_student=np.float32, output_target=np.float32, truncated=np.uint8, reset=np.uint8, episode_start=np.int32, returns=np.float32, episode_length=np.int32)
        ptr, spaces = {}, set()
        for k, d in dt.items():
            p, ms, _ = self._arg(out.get(k), d, None, k)
            ptr[k] = p
            if out.get(k) is not None:
                spaces.add(ms)
        if len(spaces) != 1:
            raise ValueError("dagger_gather: all buffers must live in the same memory space")
        o = L.DaggerOut(spaces.pop(), 0, int(out["input_student"].shape[0]), ptr["input_student"], ptr["output_target"], ptr["truncated"], ptr["reset"],
                        ptr["episode_start"], ptr["returns"], ptr["episode_length"])
        rows = ctypes.c_int64(0)
        self._check(self._lib.b200l2f_dagger_gather(self._h, n_steps, int(no_auto_reset), ctypes.byref(o), ctypes.byref(rows)))
        out = dict(out)
        out["rows"] = int(rows.value)
        return out

    def get_hidden(self):
        h = np.zeros((self.N_ENVIRONMENTS, self.policy.hidden_dim), np.float32)
        g = np.zeros(self.N_ENVIRONMENTS, np.int32)
        self._check(self._lib.b200l2f_policy_get_hidden(self._h, h.ctypes.data, g.ctypes.data, L.HOST))
        return h, g

    def set_hidden(self, hidden=None, gru_step=None):
        ph, ms, _ = self._arg(hidden, np.float32, (self.N_ENVIRONMENTS, self.policy.hidden_dim), "hidden")
        pg, msg, _ = self._arg(gru_step, np.int32, (self.N_ENVIRONMENTS,), "gru_step")
        self._check(self._lib.b200l2f_policy_set_hidden(self._h, ph, pg, ms if hidden is not None else msg))

    # ---- the fused hot path
    def rollout(self, n_steps, record=(), state_stride=1, no_auto_reset=False, out=None):
        """T closed-loop steps in ONE kernel launch, in place on slot 0.
        record: subset of {"states","observations","actions","rewards","terminated","returns","episode_length"} -> numpy arrays (host),
        or pass `out` = dict of preallocated torch CUDA tensors / numpy arrays (all in one memory space)."""
        n, T = self.N_ENVIRONMENTS, n_steps
        shapes = {"states": ((T // max(state_stride, 1) + 1, n, self.STATE_DIM), np.float32), "observations": ((T, n, 22 if self.policy is None or self.policy.arch == L.POLICY_RAPTOR_GRU else self.OBSERVATION_DIM), np.float32),
                  "actions": ((T, n, 4), np.float32), "rewards": ((T, n), np.float32), "terminated": ((T, n), np.uint8),
                  "returns": ((n,), np.float32), "episode_length": ((n,), np.int32)}
        out = dict(out) if out else {}
        for k in record:
            if k not in out:
                out[k] = np.zeros(shapes[k][0], shapes[k][1])
        ro = L.RolloutOut()
        ro.state_stride = state_stride
        spaces = set()
        for k, v in out.items():
            p, ms, _ = self._arg(v, shapes[k][1], shapes[k][0], k)
            setattr(ro, k, p)
            spaces.add(ms)
        if len(spaces) > 1:
            raise ValueError("rollout: all outputs must live in one memory space")
        ro.memspace = spaces.pop() if spaces else L.DEVICE
        self._check(self._lib.b200l2f_rollout(self._h, T, int(no_auto_reset), ctypes.byref(ro) if out else None))
        return out
