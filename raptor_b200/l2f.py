"""README-compatible `l2f` module surface (R/README.md:40-105) on top of the CUDA engine.

    import raptor_b200.l2f as l2f
    from raptor_b200.l2f import vector8 as vector        # any vectorN: l2f.vector(N)

The pip wheel `l2f==2.0.18` is not part of the reference tree; its surface is specified by the README only.  We pin it as: the default
`l2f::Specification<float, size_t>` (OBSERVATION_DIM 82, the README slices `[:, :22]` for the Raptor policy), one RNG stream per
environment seeded `seed + env index`.  Objects keep the README's shapes: VectorEnvironment owns the engine handle,
VectorParameters / VectorRng are tokens bound to it on first use, every VectorState maps to a state slot of the handle.
"""
import json
import sys
import types

import numpy as np

from . import _lib as L
from .engine import VectorEnvironment as _Engine

_MAX_SLOTS = 8


class Device:
    """l2f.Device(): placeholder for rl_tools' device object (everything runs on the CUDA device of the environment)"""
    def __init__(self, ordinal=0):
        self.ordinal = ordinal


class UI:
    def __init__(self):
        self.ns = ""


class _EnvState:
    """host snapshot of one environment's state (ui_state.states[i].position[0] += ... in the README)"""
    def __init__(self, row, H):
        self.position = row[0:3]
        self.orientation = row[3:7]
        self.linear_velocity = row[7:10]
        self.angular_velocity = row[10:13]
        self.last_action = row[13:17]
        self.force = row[20:23]
        self.torque = row[23:26]
        self.rpm = row[26:30]
        self.current_step = int(row[30])
        self.action_history = row[31:31 + 4 * H].reshape(H, 4)


def _make_vector_module(n):
    mod = types.ModuleType("raptor_b200.l2f.vector%d" % n)

    class VectorEnvironment:
        N_ENVIRONMENTS = n
        OBSERVATION_DIM = 82
        ACTION_DIM = 4
        EPISODE_STEP_LIMIT = 500

        def __init__(self, spec=L.SPEC_DEFAULT, device=0):
            self._engine = _Engine(n, spec, device=device, n_state_slots=_MAX_SLOTS)
            self.OBSERVATION_DIM = self._engine.OBSERVATION_DIM
            self._slots = 0

        def _alloc_slot(self):
            if self._slots >= _MAX_SLOTS:
                raise RuntimeError("too many live VectorState objects for one VectorEnvironment")
            self._slots += 1
            return self._slots - 1

    class VectorParameters:
        def __init__(self):
            self._env = None

    class VectorRng:
        def __init__(self):
            self._env = None
            self._seed = 0

    class VectorState:
        def __init__(self):
            self._env = None
            self._slot = None
            self._host = None   # detached host snapshot (copy.copy(state))

        def _bind(self, env):
            if self._env is None:
                self._env, self._slot = env, env._alloc_slot()
                if self._host is not None:
                    env._engine.set_state(np.ascontiguousarray(self._host), slot=self._slot)
                    self._host = None
            return self._slot

        def assign(self, other):
            if other._env is None:
                raise RuntimeError("assign: source state has never been used with an environment")
            self._bind(other._env)
            self._env._engine.copy_state(self._slot, other._slot)

        def numpy(self):
            if self._env is None:
                return self._host
            return self._env._engine.get_state(slot=self._slot)

        @property
        def states(self):
            if self._host is None:
                self._host = self.numpy()
            H = (self._host.shape[1] - 44) // 4
            return [_EnvState(self._host[i], H) for i in range(self._host.shape[0])]

        def __copy__(self):
            c = VectorState()
            c._host = self.numpy().copy()
            return c

    def initialize_rng(device, rng, seed):
        rng._seed = int(seed)
        if rng._env is not None:
            rng._env._engine.initialize_rng(rng._seed, 0)

    def _bind_rng(env, rng):
        if rng._env is None:
            rng._env = env
            env._engine.initialize_rng(rng._seed, 0)

    def initialize_environment(device, env):
        env._engine.initialize_environment()
        env._engine.initial_parameters()

    def sample_initial_parameters(device, env, params, rng):
        _bind_rng(env, rng)
        params._env = env
        env._engine.sample_initial_parameters()

    def initial_parameters(device, env, params):
        params._env = env
        env._engine.initial_parameters()

    def sample_initial_state(device, env, params, state, rng):
        _bind_rng(env, rng)
        env._engine.sample_initial_state(slot=state._bind(env))

    def initial_state(device, env, params, state):
        env._engine.initial_state(slot=state._bind(env))

    def observe(device, env, params, state, observation, rng):
        _bind_rng(env, rng)
        env._engine.observe(observation, slot=state._bind(env))

    def step(device, env, params, state, action, next_state, rng):
        _bind_rng(env, rng)
        a = np.ascontiguousarray(action, np.float32)
        dts = env._engine.step(a, slot=state._bind(env), next_slot=next_state._bind(env))
        return [float(x) for x in dts]

    # ---- UI messages (R/README.md:72-77,86-88): same channels as L2F/ui.h:37-116, minimal payloads (no websocket in this repo)
    def set_ui_message(device, env, ui):
        return json.dumps({"namespace": ui.ns, "channel": "setUI", "data": {"environments": n}})

    def set_parameters_message(device, env, params, ui):
        p = env._engine.get_parameters()
        data = [{"parameters": {"dynamics": {"mass": float(r[60]), "rotor_positions": r[0:12].reshape(4, 3).tolist()}}} for r in p]
        return json.dumps({"namespace": ui.ns, "channel": "setParameters", "data": data})

    def set_state_action_message(device, env, params, ui, state, action):
        s = state.numpy()
        a = np.asarray(action, dtype=np.float32)
        data = [{"state": {"position": s[i, 0:3].tolist(), "orientation": s[i, 3:7].tolist(), "linear_velocity": s[i, 7:10].tolist(),
                           "angular_velocity": s[i, 10:13].tolist(), "rpm": s[i, 26:30].tolist()}, "action": a[i].tolist()} for i in range(s.shape[0])]
        return json.dumps({"namespace": ui.ns, "channel": "setStateAction", "data": data})

    for k, v in dict(VectorEnvironment=VectorEnvironment, VectorParameters=VectorParameters, VectorRng=VectorRng, VectorState=VectorState,
                     initialize_rng=initialize_rng, initialize_environment=initialize_environment, sample_initial_parameters=sample_initial_parameters,
                     initial_parameters=initial_parameters, sample_initial_state=sample_initial_state, initial_state=initial_state, observe=observe, step=step,
                     set_ui_message=set_ui_message, set_parameters_message=set_parameters_message, set_state_action_message=set_state_action_message).items():
        setattr(mod, k, v)
    return mod


_cache = {}


def vector(n):
    """l2f.vector(N): the vectorN module for any N"""
    if n not in _cache:
        _cache[n] = _make_vector_module(n)
        sys.modules[_cache[n].__name__] = _cache[n]
    return _cache[n]


def __getattr__(name):   # l2f.vector8, l2f.vector64, ...
    if name.startswith("vector") and name[6:].isdigit():
        return vector(int(name[6:]))
    raise AttributeError(name)
