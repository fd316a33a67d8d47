"""The README loop (R/README.md:40-105, without UI/websocket) run verbatim against the README-compatible modules, compared with the
golden trajectory of BASELINE config 1 (default l2f spec, 8 envs x 500 steps, seed 0, Raptor checkpoint) generated from the reference."""
import os
from copy import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_readme_loop_matches_reference_config1():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200.l2f as l2f
    from raptor_b200.l2f import vector8 as vector
    from raptor_b200.foundation_policy import Raptor
    g = np.load(os.path.join(G, "default_8x500.npz"))

    policy = Raptor()
    device = l2f.Device()
    rng = vector.VectorRng()
    env = vector.VectorEnvironment()
    ui = l2f.UI()
    params = vector.VectorParameters()
    state = vector.VectorState()
    observation = np.zeros((env.N_ENVIRONMENTS, env.OBSERVATION_DIM), dtype=np.float32)
    next_state = vector.VectorState()
    vector.initialize_rng(device, rng, 0)
    vector.initialize_environment(device, env)
    vector.sample_initial_parameters(device, env, params, rng)
    vector.sample_initial_state(device, env, params, state, rng)
    np.testing.assert_allclose(state.numpy(), g["states0"], rtol=2e-6, atol=1e-7)

    ui_state = copy(state)
    for i, s in enumerate(ui_state.states):
        s.position[0] += i * 0.1
    assert '"channel": "setStateAction"' in vector.set_state_action_message(device, env, params, ui, ui_state, np.zeros((8, 4)))
    assert '"channel": "setParameters"' in vector.set_parameters_message(device, env, params, ui)

    policy.reset()
    actions = []
    for _ in range(100):
        vector.observe(device, env, params, state, observation, rng)
        action = policy.evaluate_step(observation[:, :22])
        dts = vector.step(device, env, params, state, action, next_state, rng)
        state.assign(next_state)
        actions.append(action.copy())
        assert abs(dts[-1] - 0.01) < 1e-9
    actions = np.array(actions)
    scale = np.maximum(np.abs(g["actions"][:100]).max(axis=(0, 2)), 0.1)
    assert (np.abs(actions - g["actions"][:100]).max(axis=(0, 2)) <= 1e-4 * scale).all()
    got = state.numpy()
    want = g["states"][list(g["state_steps"]).index(100)]
    for sl, floor in [(slice(0, 3), 0.1), (slice(3, 7), 1.0), (slice(7, 10), 0.1), (slice(10, 13), 0.1), (slice(26, 30), 0.1)]:
        sc = np.maximum(np.abs(want[:, sl]).max(axis=1), floor)
        assert (np.abs(got[:, sl] - want[:, sl]).max(axis=1) <= 1e-4 * sc).all()
    assert np.array_equal(got[:, 30], want[:, 30])   # ring-buffer index


def test_pybind_modules_run_the_readme_loop_identically():
    """the compiled modules (pybind11, csrc/pybind_l2f.cpp) against the ctypes ones: same calls, same bits"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200.l2f as l2f_py
    import raptor_b200._l2f_pybind as l2f_cc
    from raptor_b200.foundation_policy import Raptor as RaptorPy

    def run(l2f, Raptor):
        vector = l2f.vector8
        policy = Raptor()
        device, rng, env, ui, params, state, next_state = l2f.Device(), vector.VectorRng(), vector.VectorEnvironment(), l2f.UI(), vector.VectorParameters(), vector.VectorState(), vector.VectorState()
        observation = np.zeros((env.N_ENVIRONMENTS, env.OBSERVATION_DIM), dtype=np.float32)
        vector.initialize_rng(device, rng, 0)
        vector.initialize_environment(device, env)
        vector.sample_initial_parameters(device, env, params, rng)
        vector.sample_initial_state(device, env, params, state, rng)
        ui.ns = "ns"
        ui_state = copy(state)
        for i, s in enumerate(ui_state.states):
            s.position[0] += i * 0.1
        msg = json.loads(vector.set_state_action_message(device, env, params, ui, ui_state, np.zeros((8, 4))))
        assert msg["channel"] == "setStateAction" and abs(msg["data"][3]["state"]["position"][0] - (state.numpy()[3, 0] + 0.3)) < 1e-6
        assert json.loads(vector.set_parameters_message(device, env, params, ui))["channel"] == "setParameters"
        assert json.loads(vector.set_ui_message(device, env, ui))["namespace"] == "ns"
        policy.reset()
        acts = []
        for _ in range(60):
            vector.observe(device, env, params, state, observation, rng)
            action = policy.evaluate_step(observation[:, :22])
            dts = vector.step(device, env, params, state, action, next_state, rng)
            state.assign(next_state)
            acts.append(np.array(action))
            assert abs(dts[-1] - 0.01) < 1e-9
        return np.array(acts), state.numpy(), observation.copy()

    import json
    a1, s1, o1 = run(l2f_py, RaptorPy)
    a2, s2, o2 = run(l2f_cc, l2f_cc.foundation_policy.Raptor)
    assert np.array_equal(a1, a2) and np.array_equal(s1, s2) and np.array_equal(o1, o2)


def test_readme_script_runs_unmodified(tmp_path):
    """the reference's README script itself (tests/golden/readme_script.py = R/README.md:40-105 byte for byte, extracted by
    tests/golden/extract_readme_script.py), imports untouched: `import l2f`, `from l2f import vector8 as vector`, `from foundation_policy import Raptor`
    resolve to the repository's top-level packages.  A stand-in for the ui-server (websocket on localhost:13337, handshake as L2F/ui.h expects) records
    the messages; the last setStateAction message must carry the golden 500-step state of BASELINE config 1."""
    import asyncio
    import json
    import subprocess
    import sys
    import threading
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    websockets = pytest.importorskip("websockets")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(G, "readme_script.py")
    received = []
    ready = threading.Event()
    stop = {}

    async def handler(ws):
        await ws.send(json.dumps({"channel": "handshake", "data": {"namespace": "pytest"}}))
        try:
            async for msg in ws:
                received.append(msg)
        except Exception:
            pass

    def serve():
        loop = asyncio.new_event_loop()
        asyncio.set_event_loop(loop)

        async def run():
            stop["event"] = asyncio.Event()
            async with websockets.serve(handler, "localhost", 13337):
                ready.set()
                await stop["event"].wait()
        stop["loop"] = loop
        loop.run_until_complete(run())
    th = threading.Thread(target=serve, daemon=True)
    th.start()
    assert ready.wait(10)
    try:
        r = subprocess.run([sys.executable, script], cwd=root, capture_output=True, text=True, timeout=240, env=dict(os.environ, PYTHONPATH=root))
    finally:
        stop["loop"].call_soon_threadsafe(stop["event"].set)
        th.join(5)
    assert r.returncode == 0, r.stderr[-3000:]
    msgs = [json.loads(m) for m in received]
    chans = [m["channel"] for m in msgs]
    assert chans[0] == "setUI" and chans[1] == "setParameters" and chans.count("setStateAction") == 501 and all(m["namespace"] == "pytest" for m in msgs)
    g = np.load(os.path.join(G, "default_8x500.npz"))
    want = g["states"][list(g["state_steps"]).index(500)]
    last = msgs[-1]["data"]
    pos = np.array([d["state"]["position"] for d in last]) - np.array([[0.1 * i, 0, 0] for i in range(8)])
    np.testing.assert_allclose(pos, want[:, 0:3], rtol=0, atol=5e-3)          # 500 closed-loop steps: bounded-close (the 1e-4 bound is checked on the first 100 above)
    np.testing.assert_allclose(np.array([d["action"] for d in last]), g["actions"][499], rtol=0, atol=5e-3)
