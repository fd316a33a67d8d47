"""Pins the plain-C oracle (oracle/l2f_oracle.c) to the committed golden vectors that were generated
from the unmodified reference (tests/golden/generate.py).  Runs without /root/reference, i.e. also on
the GPU box.  glibc libm can differ between hosts by an ulp in expf/tanhf/cosf, so the closed-loop
comparisons use a tight tolerance instead of bit equality (bit equality against the reference built on
the SAME host is asserted in test_oracle_vs_reference.py)."""
import os

import numpy as np
import pytest

from oracle import binding as B

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 2e-5, 2e-6


def load(name):
    return np.load(os.path.join(G, name))


def test_checkpoint_known_answer_test(port):
    """the KAT embedded in the checkpoint (checkpoint.h:197-214), replayed like c_backend.h:54-83"""
    k = load("raptor_kat.npz")
    pol = port.make_policy(k["blob"])
    errs = []
    for b in range(2):
        h = k["h0"][None].astype(np.float32).copy()
        st = np.zeros(1, np.int32)
        for t in range(500):
            a, _, _ = port.policy_evaluate_step(pol, k["input"][t, b:b + 1], h, st)
            errs.append(np.abs(a[0] - k["output"][t, b]))
    errs = np.array(errs)
    assert errs.mean() < 5e-7 and errs.max() < 3e-6


@pytest.mark.parametrize("name", ["default_8x500.npz", "raptor_dr_64x100.npz", "raptor_noise_8x50.npz"])
def test_closed_loop_fixture(port, name):
    g = load(name)
    spec = int(g["spec"])
    pol = port.make_policy(load("raptor_kat.npz")["blob"])
    params = np.ascontiguousarray(g["params"])
    s, r = g["states0"].copy(), g["rng0"].copy()
    n = s.shape[0]
    T = g["actions"].shape[0]
    h = np.tile(load("raptor_kat.npz")["h0"], (n, 1)).astype(np.float32)
    gs = np.zeros(n, np.int32)
    out = port.rollout(spec, pol, params, s, r, T, hidden=h, gru_step=gs)
    assert np.array_equal(r, g["final_rng"])          # integer stream: bit exact
    assert np.array_equal(gs, g["final_gru_step"])
    assert np.array_equal(out["terminated"], g["terminated"])
    np.testing.assert_allclose(out["actions"], g["actions"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out["states"][g["state_steps"]], g["states"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(out["rewards"], g["rewards"], rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(h, g["final_hidden"], rtol=RTOL, atol=ATOL)


def test_dr_parameter_sampling_fixture(port):
    g = load("raptor_dr_64x100.npz")
    rng = g["rng_before_params"].copy()
    params = port.sample_initial_parameters_n(int(g["spec"]), g["env_params"], rng)
    np.testing.assert_allclose(params, g["params"], rtol=1e-6, atol=0)
    states = port.sample_initial_state_n(int(g["spec"]), params, rng)
    assert np.array_equal(rng, g["rng0"])
    np.testing.assert_allclose(states, g["states0"], rtol=1e-6, atol=1e-7)


def test_teacher_env_fixture(port):
    g = load("teacher_16x20.npz")
    spec, p = int(g["spec"]), g["params"]
    n = g["states"].shape[1]
    rng = g["rng0"].copy()
    s = g["states"][0].copy()
    for t in range(g["actions"].shape[0]):
        for i in range(n):
            r = rng[i:i + 1].copy()
            o = port.observe(spec, p, s[i], r)
            nx, _ = port.step(spec, p, s[i], g["actions"][t, i], r)
            np.testing.assert_allclose(o, g["observations"][t, i], rtol=RTOL, atol=ATOL)
            np.testing.assert_allclose(nx, g["states"][t + 1, i], rtol=RTOL, atol=ATOL)
            assert abs(port.reward(spec, p, s[i], g["actions"][t, i], nx) - g["rewards"][t, i]) < 1e-4
            assert port.terminated(spec, p, nx) == bool(g["terminated"][t, i])
            s[i] = g["states"][t + 1, i]   # re-anchor on the fixture: one-step comparison
            rng[i] = r[0]
    assert np.array_equal(rng, g["final_rng"])
