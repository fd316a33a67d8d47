"""README-compatible `foundation_policy` module surface (R/README.md:19-24,94-97): Raptor().reset() / .evaluate_step(obs[N,22]) -> [N,4].

The actor is the reference checkpoint (Dense 22->16 ReLU -> GRU 16 -> Dense 16->4, 2084 parameters) evaluated by the CUDA engine.
The pip wheel `foundation-policy==1.0.1` is not in the reference tree; unpinned details are fixed as: Mode<Evaluation>, hidden state
auto-reset every 500 steps as in training (rl_tools/nn/layers/gru/operations_generic.h:80,403) unless `no_auto_reset=True`
(the deployment executor's NoAutoResetMode, rl_tools/inference/executor/operations_generic.h:179)."""
import numpy as np

from . import _lib as L
from .engine import Checkpoint, VectorEnvironment, raptor_policy_blob


class Raptor:
    def __init__(self, device=0, no_auto_reset=False, checkpoint=None):
        """checkpoint: path of an rl-tools `checkpoint.h` code export or `checkpoint.h5` with a Dense-GRU-Dense actor (another training run of
        the foundation policy); default = the weights of the reference's published checkpoint shipped in raptor_b200/data"""
        self._policy = dict(blob=raptor_policy_blob()) if checkpoint is None else Checkpoint(path=checkpoint).policy_kwargs()
        self._device = device
        self._no_auto_reset = no_auto_reset
        self._engine = None
        self._pending_reset = True

    def _ensure(self, n):
        if self._engine is None or self._engine.N_ENVIRONMENTS != n:
            self._engine = VectorEnvironment(n, L.SPEC_RAPTOR, device=self._device)
            self._engine.load_policy(**self._policy)
            self._pending_reset = False   # load_policy resets

    def reset(self):
        if self._engine is None:
            self._pending_reset = True
        else:
            self._engine.policy_reset()

    def evaluate_step(self, observation):
        obs = np.ascontiguousarray(observation, np.float32)
        if obs.ndim != 2 or obs.shape[1] < 22:
            raise ValueError("Raptor.evaluate_step expects [batch, 22] observations (position, rotation matrix, linear velocity, angular velocity, previous action)")
        self._ensure(obs.shape[0])
        if self._pending_reset:
            self._engine.policy_reset()
            self._pending_reset = False
        return self._engine.policy_evaluate_step(obs, no_auto_reset=self._no_auto_reset)
