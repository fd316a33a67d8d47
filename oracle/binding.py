"""oracle/binding.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes bindings for the two CPU checkers:

* ``Port``  -> oracle/libl2f_oracle.so      (plain-C restatement, oracle/l2f_oracle.c)
* ``Ref``   -> oracle/_ref/libl2f_ref.so    (the unmodified reference compiled from /root/reference
                                             by oracle/Makefile; present only if it was built in the
                                             build container -- it travels to the GPU box as a file)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  Nothing under raptor_b200/ does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS_DIM = 145
SPEC_DEFAULT, SPEC_DEFAULT_DR, SPEC_RAPTOR, SPEC_TEACHER, SPEC_RAPTOR_DR, SPEC_TEACHER_DR = range(6)

c_int = ctypes.c_int
c_float = ctypes.c_float
c_u64 = ctypes.c_uint64
f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
vp = ctypes.c_void_p


def build(target="all"):
    """Compile the checkers (building the checker is not using it)."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(vp)


class _Common:
    prefix = ""

    def __init__(self, path):
        self.path = path
        self.lib = ctypes.CDLL(path)
        L, P = self.lib, self.prefix
        g = lambda n: getattr(L, P + n)
        g("rng_init").restype = c_u64
        g("rng_init").argtypes = [c_u64]
        g("rng_uniform").restype = c_float
        g("rng_uniform").argtypes = [u64, c_float, c_float]
        g("rng_normal").restype = c_float
        g("rng_normal").argtypes = [u64, c_float, c_float]
        g("nominal_parameters").argtypes = [c_int, f32]
        g("sample_initial_parameters").argtypes = [c_int, f32, u64, f32]
        g("initial_state").argtypes = [c_int, f32, f32]
        g("sample_initial_state").argtypes = [c_int, f32, u64, f32]
        g("observe").argtypes = [c_int, f32, f32, u64, f32]
        g("step").restype = c_float
        g("step").argtypes = [c_int, f32, f32, f32, u64, f32]
        g("reward").restype = c_float
        g("reward").argtypes = [c_int, f32, f32, f32, f32, u64]
        g("terminated").argtypes = [c_int, f32, f32]
        self._g = g

    # ---- sizes
    def state_dim(self, spec):
        return self._g("state_dim")(spec)

    def observation_dim(self, spec):
        return self._g("observation_dim")(spec)

    def action_history_length(self, spec):
        return self._g("action_history_length")(spec)

    # ---- rng
    def rng_init(self, seed):
        return int(self._g("rng_init")(int(seed)))

    def rng_states(self, base_seed, n, first_env=0, warmup=0):
        """per-environment streams: seed = base_seed + global env id, optionally advanced by `warmup` uniform draws"""
        st = np.array([self.rng_init(base_seed + first_env + i) for i in range(n)], dtype=np.uint64)
        if warmup:
            for i in range(n):
                s = st[i:i + 1].copy()
                for _ in range(warmup):
                    self._g("rng_uniform")(s, 0.0, 1.0)
                st[i] = s[0]
        return st

    def rng_uniform(self, state, lo, hi):
        return float(self._g("rng_uniform")(state, lo, hi))

    def rng_normal(self, state, mean, std):
        return float(self._g("rng_normal")(state, mean, std))

    # ---- env (single environment, flat rows)
    def nominal_parameters(self, spec):
        p = np.zeros(PARAMS_DIM, np.float32)
        self._g("nominal_parameters")(spec, p)
        return p

    def sample_initial_parameters(self, spec, env_p, rng):
        out = np.zeros(PARAMS_DIM, np.float32)
        self._g("sample_initial_parameters")(spec, np.ascontiguousarray(env_p, np.float32), rng, out)
        return out

    def initial_state(self, spec, p):
        s = np.zeros(self.state_dim(spec), np.float32)
        self._g("initial_state")(spec, np.ascontiguousarray(p, np.float32), s)
        return s

    def sample_initial_state(self, spec, p, rng):
        s = np.zeros(self.state_dim(spec), np.float32)
        self._g("sample_initial_state")(spec, np.ascontiguousarray(p, np.float32), rng, s)
        return s

    def observe(self, spec, p, s, rng):
        o = np.zeros(self.observation_dim(spec), np.float32)
        self._g("observe")(spec, np.ascontiguousarray(p, np.float32), np.ascontiguousarray(s, np.float32), rng, o)
        return o

    def step(self, spec, p, s, a, rng):
        n = np.zeros(self.state_dim(spec), np.float32)
        dt = self._g("step")(spec, np.ascontiguousarray(p, np.float32), np.ascontiguousarray(s, np.float32), np.ascontiguousarray(a, np.float32), rng, n)
        return n, float(dt)

    def reward(self, spec, p, s, a, n, rng=None):
        rng = np.zeros(1, np.uint64) if rng is None else rng
        return float(self._g("reward")(spec, np.ascontiguousarray(p, np.float32), np.ascontiguousarray(s, np.float32), np.ascontiguousarray(a, np.float32), np.ascontiguousarray(n, np.float32), rng))

    def terminated(self, spec, p, s):
        return bool(self._g("terminated")(spec, np.ascontiguousarray(p, np.float32), np.ascontiguousarray(s, np.float32)))

    # ---- vector helpers (loops over the single-env entry points)
    def sample_initial_parameters_n(self, spec, env_p, rngs):
        n = len(rngs)
        out = np.zeros((n, PARAMS_DIM), np.float32)
        for i in range(n):
            r = rngs[i:i + 1].copy()
            out[i] = self.sample_initial_parameters(spec, env_p, r)
            rngs[i] = r[0]
        return out

    def sample_initial_state_n(self, spec, params, rngs):
        n = len(rngs)
        out = np.zeros((n, self.state_dim(spec)), np.float32)
        for i in range(n):
            r = rngs[i:i + 1].copy()
            out[i] = self.sample_initial_state(spec, params[i], r)
            rngs[i] = r[0]
        return out


class OraclePolicy(ctypes.Structure):
    _fields_ = [("arch", c_int), ("input_dim", c_int), ("hidden_dim", c_int), ("output_dim", c_int),
                ("standardize", c_int), ("head", c_int), ("blob", ctypes.POINTER(c_float))]


POLICY_RAPTOR_GRU, POLICY_MLP = 0, 1
HEAD_IDENTITY, HEAD_SQUASH_EVAL, HEAD_PPO_GAUSSIAN, HEAD_SQUASH_SAMPLE = 0, 1, 2, 3


def new_sequential_batch(B, L, obs):
    """SequentialBatch tensors (off_policy_runner.h:96-141) for SEQUENCE_LENGTH L: [L + 1] = the padded (`*_base`) tensors"""
    return dict(observations_actions=np.zeros((L + 1, B, obs + 4), np.float32), rewards=np.zeros((L, B), np.float32), terminated=np.zeros((L, B), np.uint8),
                reset=np.zeros((L, B), np.uint8), next_reset=np.zeros((L + 1, B), np.uint8), final_step_mask=np.zeros((L, B), np.uint8),
                next_final_step_mask=np.zeros((L + 1, B), np.uint8), env_index=np.zeros(B, np.int32), sample_index=np.zeros(B, np.int32))


def synthetic_replay_rings(rs, n, capacity, obs, p_truncated=0.12, fill=None):
    """replay rings with arbitrary contents but CONSISTENT bookkeeping (what `add` maintains, replay_buffer/operations_generic.h:54-79): rows written in
    order from row 0, `truncated` ends an episode, episode_start[row] = first row of the row's episode; fill[e] rows were added to ring e (>= capacity: wrapped)"""
    D = 2 * obs + 7
    runner = new_off_policy_runner(n, capacity, obs)
    for e in range(n):
        total = int(fill[e]) if fill is not None else int(rs.randint(capacity // 2, 3 * capacity))
        pos, start = 0, 0
        for _ in range(total):
            row = rs.uniform(-1, 1, D).astype(np.float32)
            trunc = rs.uniform() < p_truncated
            row[D - 2] = float(trunc and rs.uniform() < 0.5); row[D - 1] = float(trunc)
            runner["replay"][e, pos] = row
            runner["episode_start"][e, pos] = start
            pos = (pos + 1) % capacity
            if trunc:
                start = pos
            if pos == 0:
                runner["full"][e] = 1
        runner["position"][e] = pos
        runner["current_episode_start"][e] = start
    return runner


def new_off_policy_runner(n, capacity, obs_dim):
    """runner state after rl_tools::init(device, runner): everything truncated, empty replay rings (off_policy_runner/operations_generic.h:163-180)"""
    return dict(episode_step=np.zeros(n, np.int32), episode_return=np.zeros(n, np.float32), truncated=np.ones(n, np.uint8),
                replay=np.zeros((n, capacity, 2 * obs_dim + 7), np.float32), episode_start=np.zeros((n, capacity), np.int32),
                position=np.zeros(n, np.int32), full=np.zeros(n, np.uint8), current_episode_start=np.zeros(n, np.int32))


class Port(_Common):
    """plain-C restatement (oracle/l2f_oracle.c)"""
    prefix = "oracle_"
    kind = "port"

    def __init__(self, fast=False):
        name = "libl2f_oracle_fast.so" if fast else "libl2f_oracle.so"
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            build("port")
        super().__init__(path)
        L = self.lib
        L.oracle_policy_num_parameters.argtypes = [ctypes.POINTER(OraclePolicy)]
        L.oracle_policy_evaluate_step.argtypes = [ctypes.POINTER(OraclePolicy), c_int, f32, c_int, vp, vp, c_int, vp, f32, vp, vp]
        L.oracle_rollout.argtypes = [c_int, ctypes.POINTER(OraclePolicy), c_int, c_int, c_int, f32, f32, u64, vp, vp, c_int, vp, vp, vp, vp, vp]
        L.oracle_collect.argtypes = [c_int, ctypes.POINTER(OraclePolicy), c_int, c_int, c_int, c_int, f32, f32, f32, u64, vp, vp, vp, f32, c_int]
        L.oracle_off_policy_steps.argtypes = [c_int, ctypes.POINTER(OraclePolicy), c_int, c_int, c_int, c_int, c_int, f32, f32, f32, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.oracle_gather_batch.argtypes = [c_int, c_int, c_int, c_int, c_int, f32, vp, vp, c_int, u64, f32, f32, vp, vp, vp, vp, vp, vp, vp]
        L.oracle_gather_batch_sequential.argtypes = [c_int, c_int, c_int, c_int, c_int, f32, vp, vp, vp, c_int, c_int, c_int, c_int, c_int, c_float, c_int, u64, f32, f32, vp, vp, vp, vp, vp, vp, vp]
        L.oracle_evaluate_values.argtypes = [ctypes.POINTER(OraclePolicy), c_int, c_int, f32, c_int]
        L.oracle_estimate_generalized_advantages.argtypes = [c_int, c_int, f32, c_int, c_float, c_float, c_int]
        L.oracle_normalizer_update.argtypes = [c_int, c_int, f32, c_int, f32, f32, ctypes.POINTER(c_int)]
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
        i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
        L.oracle_dagger_add_to_dataset.restype = ctypes.c_longlong
        L.oracle_dagger_add_to_dataset.argtypes = [c_int, c_int, c_int, f32, f32, u8, u64, f32, f32, i32, f32, f32, u8, u8]

    def make_policy(self, blob, arch=POLICY_RAPTOR_GRU, input_dim=22, hidden_dim=16, output_dim=4, standardize=0, head=HEAD_IDENTITY):
        blob = np.ascontiguousarray(blob, np.float32)
        pol = OraclePolicy(arch, input_dim, hidden_dim, output_dim, standardize, head, blob.ctypes.data_as(ctypes.POINTER(c_float)))
        pol._keep = blob
        n = self.lib.oracle_policy_num_parameters(ctypes.byref(pol))
        assert n == blob.size, (n, blob.size)
        return pol

    def policy_evaluate_step(self, pol, obs, hidden=None, gru_step=None, no_auto_reset=False, rng=None):
        obs = np.ascontiguousarray(obs, np.float32)
        n = obs.shape[0]
        adim = pol.output_dim // 2 if pol.head in (HEAD_SQUASH_EVAL, HEAD_SQUASH_SAMPLE) else pol.output_dim
        act = np.zeros((n, adim), np.float32)
        mean = np.zeros((n, adim), np.float32)
        lp = np.zeros(n, np.float32)
        self.lib.oracle_policy_evaluate_step(ctypes.byref(pol), n, obs, obs.shape[1], _ptr(hidden), _ptr(gru_step), int(no_auto_reset), _ptr(rng), act, _ptr(mean), _ptr(lp))
        return act, mean, lp

    def rollout(self, spec, pol, params, states, rngs, T, hidden=None, gru_step=None, no_auto_reset=False, threads=1, record=True):
        """runs T closed-loop steps IN PLACE on states/rngs/hidden/gru_step; returns dict of recorded arrays"""
        n = states.shape[0]
        sd, od = self.state_dim(spec), self.observation_dim(spec)
        out = {}
        if record:
            out = dict(states=np.zeros((T + 1, n, sd), np.float32), observations=np.zeros((T, n, od), np.float32),
                       actions=np.zeros((T, n, 4), np.float32), rewards=np.zeros((T, n), np.float32), terminated=np.zeros((T, n), np.uint8))
        self.lib.oracle_rollout(spec, ctypes.byref(pol), n, T, threads, params, states, rngs, _ptr(hidden), _ptr(gru_step), int(no_auto_reset),
                                _ptr(out.get("states")), _ptr(out.get("observations")), _ptr(out.get("actions")), _ptr(out.get("rewards")), _ptr(out.get("terminated")))
        return out

    def collect(self, spec, pol, env_params, params, states, rngs, episode_step, episode_return, truncated, T, step_limit, threads=1):
        n = states.shape[0]
        D = self.observation_dim(spec) + 15
        data = np.zeros(((T + 1) * n, D), np.float32)
        self.lib.oracle_collect(spec, ctypes.byref(pol), n, T, threads, step_limit, np.ascontiguousarray(env_params, np.float32), params, states, rngs,
                                _ptr(episode_step), _ptr(episode_return), _ptr(truncated), data, D)
        return data

    def off_policy_steps(self, spec, pol, env_params, params, states, rngs, runner, T, step_limit, sample_parameters=True, with_states=False):
        """T runner steps IN PLACE on params/states/rngs and the runner dict (episode_step, episode_return, truncated, replay [n, capacity, 2*OBS+7],
        episode_start [n, capacity], position, full, current_episode_start [n]); optional states / next_states [n, capacity, SD]"""
        n, capacity = runner["replay"].shape[:2]
        if with_states and "states" not in runner:
            runner["states"] = np.zeros((n, capacity, self.state_dim(spec)), np.float32)
            runner["next_states"] = np.zeros((n, capacity, self.state_dim(spec)), np.float32)
        self.lib.oracle_off_policy_steps(spec, ctypes.byref(pol), n, T, step_limit, capacity, int(sample_parameters), np.ascontiguousarray(env_params, np.float32), params, states, rngs,
                                         _ptr(runner["episode_step"]), _ptr(runner["episode_return"]), _ptr(runner["truncated"]), _ptr(runner["replay"]), _ptr(runner["episode_start"]),
                                         _ptr(runner["position"]), _ptr(runner["full"]), _ptr(runner["current_episode_start"]), _ptr(runner.get("states")), _ptr(runner.get("next_states")))
        return runner

    def gather_batch(self, runner, rngs, max_episode_length, env_begin=0, env_count=None):
        """SEQUENCE_LENGTH-1 batch from the replay rings of `runner`, one RNG stream per sample (rngs [B], advanced in place)"""
        n, capacity, D = runner["replay"].shape
        obs = (D - 7) // 2
        B = rngs.shape[0]
        out = dict(observations_actions=np.zeros((2, B, obs + 4), np.float32), rewards=np.zeros(B, np.float32), terminated=np.zeros(B, np.uint8), reset=np.zeros(B, np.uint8),
                   next_reset=np.zeros((2, B), np.uint8), final_step_mask=np.zeros(B, np.uint8), next_final_step_mask=np.zeros((2, B), np.uint8),
                   env_index=np.zeros(B, np.int32), sample_index=np.zeros(B, np.int32))
        self.lib.oracle_gather_batch(obs, capacity, max_episode_length, env_begin, n if env_count is None else env_count, runner["replay"], _ptr(runner["position"]), _ptr(runner["full"]),
                                     B, rngs, out["observations_actions"], out["rewards"], *[_ptr(out[k]) for k in ("terminated", "reset", "next_reset", "final_step_mask", "next_final_step_mask", "env_index", "sample_index")])
        return out

    def gather_batch_sequential(self, runner, rngs, max_episode_length, sequence_length, include_first_step_in_targets=True, always_sample_from_initial_state=True,
                                random_seq_length=True, enable_nominal_sequence_length_probability=True, nominal_sequence_length_probability=0.5, env_begin=0, env_count=None):
        """batch of sequences (any SEQUENCE_LENGTH) from the replay rings of `runner`, one RNG stream per sample (rngs [B], advanced in place); the flag defaults
        are the reference's for SEQUENCE_LENGTH > 1 (off_policy_runner.h:78-85)"""
        n, capacity, D = runner["replay"].shape
        obs = (D - 7) // 2
        B, L = rngs.shape[0], sequence_length
        out = new_sequential_batch(B, L, obs)
        self.lib.oracle_gather_batch_sequential(obs, capacity, max_episode_length, env_begin, n if env_count is None else env_count, runner["replay"], _ptr(runner["episode_start"]),
                                                _ptr(runner["position"]), _ptr(runner["full"]), L, int(include_first_step_in_targets), int(always_sample_from_initial_state),
                                                int(random_seq_length), int(enable_nominal_sequence_length_probability), nominal_sequence_length_probability, B, rngs,
                                                out["observations_actions"], out["rewards"],
                                                *[_ptr(out[k]) for k in ("terminated", "reset", "next_reset", "final_step_mask", "next_final_step_mask", "env_index", "sample_index")])
        return out

    def evaluate_values(self, critic, data, n, T):
        """critic values over all (T+1)*n observation rows -> the all_values column, in place"""
        self.lib.oracle_evaluate_values(ctypes.byref(critic), n, T, data, data.shape[1])

    def estimate_generalized_advantages(self, data, n, T, gamma=0.99, lam=0.95, ignore_termination=False):
        self.lib.oracle_estimate_generalized_advantages(n, T, data, data.shape[1], gamma, lam, int(ignore_termination))

    def normalizer_update(self, data, n, T, mean, std, age):
        a = c_int(age)
        self.lib.oracle_normalizer_update(n, T, data, data.shape[1], mean, std, ctypes.byref(a))
        return int(a.value)

    def dagger_add_to_dataset(self, params, states, terminated, rng, teacher_blobs, offsets, episodes_per_teacher):
        """states [T, n, SD] step-major, terminated [T, n]; returns dict(rows, episode_start, input_student, output_target, truncated, reset)"""
        T, n = terminated.shape
        cap = T * n
        out = dict(episode_start=np.zeros(cap, np.int32), input_student=np.zeros((cap, 22), np.float32), output_target=np.zeros((cap, 4), np.float32),
                   truncated=np.zeros(cap, np.uint8), reset=np.zeros(cap, np.uint8))
        out["rows"] = int(self.lib.oracle_dagger_add_to_dataset(n, T, episodes_per_teacher, np.ascontiguousarray(params, np.float32), np.ascontiguousarray(states, np.float32),
                                                                np.ascontiguousarray(terminated, np.uint8), rng, np.ascontiguousarray(teacher_blobs, np.float32),
                                                                np.ascontiguousarray(offsets, np.float32), out["episode_start"], out["input_student"], out["output_target"],
                                                                out["truncated"], out["reset"]))
        return out

    def hardware_threads(self):
        return int(self.lib.oracle_hardware_threads())


class Ref(_Common):
    """the unmodified reference (oracle/ref_l2f.cpp over /root/reference headers)"""
    prefix = "ref_"
    kind = "reference"

    @staticmethod
    def available(fast=False):
        return os.path.exists(os.path.join(HERE, "_ref", "libl2f_ref_fast.so" if fast else "libl2f_ref.so"))

    def __init__(self, fast=False):
        super().__init__(os.path.join(HERE, "_ref", "libl2f_ref_fast.so" if fast else "libl2f_ref.so"))
        L = self.lib
        L.ref_policy_kat.restype = c_float
        L.ref_policy_kat.argtypes = [ctypes.POINTER(c_float)]
        L.ref_policy_export.argtypes = [f32]
        L.ref_policy_kat_export.argtypes = [f32, f32]
        L.ref_policy_evaluate_step.argtypes = [c_int, f32, f32, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), c_int, f32]
        L.ref_policy_initial_hidden.argtypes = [f32]
        L.ref_rollout.argtypes = [c_int, c_int, c_int, c_int, f32, f32, u64, vp, vp, c_int, vp, vp, vp, vp, vp]
        L.ref_checkpoint_name.restype = ctypes.c_char_p
        if hasattr(L, "ref_collect"):
            L.ref_ppo_gamma.restype = c_float
            L.ref_ppo_lambda.restype = c_float
            L.ref_mlp_evaluate.argtypes = [c_int, c_int, f32, c_int, c_int, f32, c_int, f32, c_int]
            L.ref_collect.argtypes = [c_int, f32, c_int, f32, f32, f32, u64, vp, vp, vp, f32]
            L.ref_gae.argtypes = [c_int, f32, c_int]
            L.ref_normalizer_update.argtypes = [c_int, f32, f32, f32, ctypes.POINTER(c_int)]
        if hasattr(L, "ref_off_policy_steps"):
            L.ref_off_policy_steps.argtypes = [c_int, c_int, f32, f32, f32, f32, u64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        if hasattr(L, "ref_gather_batch"):
            L.ref_gather_batch.argtypes = [f32, vp, vp, u64, f32, f32, vp, vp, vp, vp, vp]
        if hasattr(L, "ref_gather_batch_sequential"):
            L.ref_gather_batch_sequential.argtypes = [c_int, f32, vp, vp, vp, u64, f32, f32, vp, vp, vp, vp, vp]
        if hasattr(L, "ref_dagger_add_to_dataset"):
            u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
            i32 = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
            L.ref_dagger_add_to_dataset.argtypes = [f32, f32, u8, f32, f32, i32, f32, f32, u8, u8]

    def policy_kat(self):
        mx = c_float()
        mean = self.lib.ref_policy_kat(ctypes.byref(mx))
        return float(mean), float(mx.value)

    def policy_export(self):
        blob = np.zeros(self.lib.ref_policy_num_parameters(), np.float32)
        self.lib.ref_policy_export(blob)
        return blob

    def policy_kat_export(self):
        i = np.zeros((500, 2, 22), np.float32)
        o = np.zeros((500, 2, 4), np.float32)
        self.lib.ref_policy_kat_export(i, o)
        return i, o

    def policy_evaluate_step(self, obs22, hidden, gru_step, no_auto_reset=False):
        obs22 = np.ascontiguousarray(obs22, np.float32)
        act = np.zeros((obs22.shape[0], 4), np.float32)
        self.lib.ref_policy_evaluate_step(obs22.shape[0], obs22, hidden, gru_step, int(no_auto_reset), act)
        return act

    def policy_initial_hidden(self):
        h = np.zeros(16, np.float32)
        self.lib.ref_policy_initial_hidden(h)
        return h

    def rollout(self, spec, params, states, rngs, T, hidden=None, gru_step=None, no_auto_reset=False, threads=1, record=True):
        n = states.shape[0]
        sd, od = self.state_dim(spec), self.observation_dim(spec)
        out = {}
        if record:
            out = dict(states=np.zeros((T + 1, n, sd), np.float32), observations=np.zeros((T, n, od), np.float32),
                       actions=np.zeros((T, n, 4), np.float32), rewards=np.zeros((T, n), np.float32), terminated=np.zeros((T, n), np.uint8))
        self.lib.ref_rollout(spec, n, T, threads, params, states, rngs, _ptr(hidden), _ptr(gru_step), int(no_auto_reset),
                             _ptr(out.get("states")), _ptr(out.get("observations")), _ptr(out.get("actions")), _ptr(out.get("rewards")), _ptr(out.get("terminated")))
        return out

    # ---- PPO data path (sizes fixed at compile time in oracle/ref_l2f.cpp)
    def ppo_sizes(self):
        n, t, lim = c_int(), c_int(), c_int()
        self.lib.ref_ppo_sizes(ctypes.byref(n), ctypes.byref(t), ctypes.byref(lim))
        return int(n.value), int(t.value), int(lim.value)

    def ppo_gamma_lambda(self):
        return float(self.lib.ref_ppo_gamma()), float(self.lib.ref_ppo_lambda())

    def mlp_evaluate(self, blob, in_dim, out_dim, standardize, x):
        x = np.ascontiguousarray(x, np.float32)
        y = np.zeros((x.shape[0], out_dim), np.float32)
        rc = self.lib.ref_mlp_evaluate(in_dim, out_dim, np.ascontiguousarray(blob, np.float32), int(standardize), x.shape[0], x, x.shape[1], y, out_dim)
        assert rc == 0, "ref_mlp_evaluate: shape not instantiated"
        return y

    def collect(self, spec, blob, standardize, env_params, params, states, rngs, episode_step, episode_return, truncated):
        n, T, _ = self.ppo_sizes()
        assert states.shape[0] == n
        D = self.observation_dim(spec) + 15
        data = np.zeros(((T + 1) * n, D), np.float32)
        self.lib.ref_collect(spec, np.ascontiguousarray(blob, np.float32), int(standardize), np.ascontiguousarray(env_params, np.float32), params, states, rngs,
                             _ptr(episode_step), _ptr(episode_return), _ptr(truncated), data)
        return data

    def estimate_generalized_advantages(self, spec, data, ignore_termination=False):
        self.lib.ref_gae(spec, data, int(ignore_termination))

    def normalizer_update(self, spec, data, mean, std, age):
        a = c_int(age)
        self.lib.ref_normalizer_update(spec, data, mean, std, ctypes.byref(a))
        return int(a.value)

    # ---- off-policy runner (the reference's own prologue_per_env / epilogue_per_env / replay add); fixed sizes, see off_policy_sizes()
    def off_policy_sizes(self):
        v = [c_int() for _ in range(4)]
        self.lib.ref_off_policy_sizes(*[ctypes.byref(x) for x in v])
        return tuple(int(x.value) for x in v)      # n, steps, step_limit, capacity

    def off_policy_steps(self, spec, blob, env_params, params, states, rngs, runner, sample_parameters=True, with_states=False):
        n, T, _, capacity = self.off_policy_sizes()
        assert runner["replay"].shape[:2] == (n, capacity)
        if with_states and "states" not in runner:
            runner["states"] = np.zeros((n, capacity, self.state_dim(spec)), np.float32)
            runner["next_states"] = np.zeros((n, capacity, self.state_dim(spec)), np.float32)
        rc = self.lib.ref_off_policy_steps(spec, int(sample_parameters), np.ascontiguousarray(blob, np.float32), np.ascontiguousarray(env_params, np.float32), params, states, rngs,
                                           _ptr(runner["episode_step"]), _ptr(runner["episode_return"]), _ptr(runner["truncated"]), _ptr(runner["replay"]), _ptr(runner["episode_start"]),
                                           _ptr(runner["position"]), _ptr(runner["full"]), _ptr(runner["current_episode_start"]), _ptr(runner.get("states")), _ptr(runner.get("next_states")))
        assert rc == 0, "ref_off_policy_steps: spec not instantiated"
        return runner

    def gather_batch(self, runner, rngs):
        """the reference's gather_batch_step on its own SequentialBatch (SEQUENCE_LENGTH 1, batch ref_gather_batch_size()) over the rings of `runner`"""
        B = self.lib.ref_gather_batch_size()
        assert rngs.shape[0] == B
        obs = (runner["replay"].shape[2] - 7) // 2
        out = dict(observations_actions=np.zeros((2, B, obs + 4), np.float32), rewards=np.zeros(B, np.float32), terminated=np.zeros(B, np.uint8), reset=np.zeros(B, np.uint8),
                   next_reset=np.zeros((2, B), np.uint8), final_step_mask=np.zeros(B, np.uint8), next_final_step_mask=np.zeros((2, B), np.uint8))
        self.lib.ref_gather_batch(runner["replay"], _ptr(runner["position"]), _ptr(runner["full"]), rngs, out["observations_actions"], out["rewards"],
                                  *[_ptr(out[k]) for k in ("terminated", "reset", "next_reset", "final_step_mask", "next_final_step_mask")])
        return out

    def gather_batch_sequential_configs(self):
        """the parameter sets ref_gather_batch_sequential was instantiated for: list of dicts (sequence_length, flags, probability, ring capacity)"""
        out = []
        for i in range(self.lib.ref_gather_batch_sequential_configs()):
            v = [c_int() for _ in range(5)]; prob = c_float(); cap = c_int()
            self.lib.ref_gather_batch_sequential_config(i, *[ctypes.byref(x) for x in v], ctypes.byref(prob), ctypes.byref(cap))
            out.append(dict(config=i, sequence_length=v[0].value, include_first_step_in_targets=bool(v[1].value), always_sample_from_initial_state=bool(v[2].value),
                            random_seq_length=bool(v[3].value), enable_nominal_sequence_length_probability=bool(v[4].value),
                            nominal_sequence_length_probability=prob.value, capacity=cap.value))
        return out

    def gather_batch_sequential(self, config, runner, rngs):
        """the reference's gather_batch_step for compiled parameter set `config` (batch ref_gather_batch_size()) over the rings of `runner`"""
        B = self.lib.ref_gather_batch_size()
        cfg = self.gather_batch_sequential_configs()[config]
        assert rngs.shape[0] == B and runner["replay"].shape[1] == cfg["capacity"]
        obs = (runner["replay"].shape[2] - 7) // 2
        out = new_sequential_batch(B, cfg["sequence_length"], obs)
        rc = self.lib.ref_gather_batch_sequential(config, runner["replay"], _ptr(runner["episode_start"]), _ptr(runner["position"]), _ptr(runner["full"]), rngs,
                                                  out["observations_actions"], out["rewards"], *[_ptr(out[k]) for k in ("terminated", "reset", "next_reset", "final_step_mask", "next_final_step_mask")])
        assert rc == 0
        del out["env_index"], out["sample_index"]
        return out

    # ---- checkpoint code export: the reference's own save_code (kind 1 = SAC teacher MLP + sample_and_squash, 2 = PPO standardize + MLP + log_std)
    def save_code(self, kind, blob, has_std=0, name="fixture"):
        blob = np.ascontiguousarray(blob, np.float32)
        cap = 1 << 22
        buf = ctypes.create_string_buffer(cap)
        n = self.lib.ref_save_code(kind, blob.ctypes.data_as(ctypes.POINTER(c_float)), has_std, name.encode(), buf, cap)
        if n <= 0:
            raise RuntimeError("ref_save_code failed (%d)" % n)
        return buf.raw[:n].decode()

    # ---- JSON wire format (the reference's own json / from_json; needs the nlohmann header at build time)
    def json_available(self):
        return hasattr(self.lib, "ref_json_available") and bool(self.lib.ref_json_available())

    def parameters_to_json(self, spec, row):
        row = np.ascontiguousarray(row, np.float32)
        buf = ctypes.create_string_buffer(1 << 16)
        n = self.lib.ref_parameters_to_json(spec, row.ctypes.data_as(ctypes.POINTER(c_float)), buf, 1 << 16)
        assert n > 0
        return buf.value.decode()

    def parameters_from_json(self, spec, text, row):
        out = np.array(row, np.float32, copy=True)
        self.lib.ref_parameters_from_json(spec, text.encode(), out.ctypes.data_as(ctypes.POINTER(c_float)))
        return out

    def state_to_json(self, spec, prow, srow):
        prow, srow = np.ascontiguousarray(prow, np.float32), np.ascontiguousarray(srow, np.float32)
        buf = ctypes.create_string_buffer(1 << 16)
        fp = ctypes.POINTER(c_float)
        n = self.lib.ref_state_to_json(spec, prow.ctypes.data_as(fp), srow.ctypes.data_as(fp), buf, 1 << 16)
        assert n > 0
        return buf.value.decode()

    def state_from_json(self, spec, prow, text, srow):
        out = np.array(srow, np.float32, copy=True)
        fp = ctypes.POINTER(c_float)
        self.lib.ref_state_from_json(spec, np.ascontiguousarray(prow, np.float32).ctypes.data_as(fp), text.encode(), out.ctypes.data_as(fp))
        return out

    def dagger_sizes(self):
        a, b = c_int(), c_int()
        self.lib.ref_dagger_sizes(ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def dagger_add_to_dataset(self, params, states_episode_major, terminated_episode_major, teacher_blob, offset):
        """the reference's add_to_dataset for ONE teacher: params [10, 145], states [10, 500, 48], terminated [10, 500]"""
        ne, T = self.dagger_sizes()
        cap = ne * T
        out = dict(episode_start=np.zeros(cap, np.int32), input_student=np.zeros((cap, 22), np.float32), output_target=np.zeros((cap, 4), np.float32),
                   truncated=np.zeros(cap, np.uint8), reset=np.zeros(cap, np.uint8))
        out["rows"] = int(self.lib.ref_dagger_add_to_dataset(np.ascontiguousarray(params, np.float32), np.ascontiguousarray(states_episode_major, np.float32),
                                                             np.ascontiguousarray(terminated_episode_major, np.uint8), np.ascontiguousarray(teacher_blob, np.float32),
                                                             np.ascontiguousarray(offset, np.float32), out["episode_start"], out["input_student"], out["output_target"],
                                                             out["truncated"], out["reset"]))
        return out

    def hardware_threads(self):
        return int(self.lib.ref_hardware_threads())
