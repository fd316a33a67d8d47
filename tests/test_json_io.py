"""Parameter / state JSON wire format: the engine's host-side import / export against the reference's own json / from_json
(rl/environments/l2f/operations_cpu.h) and against golden dynamics-parameter files of the reference's foundation-policy data set
(tests/golden/dynamics_parameters/*.json, copied out of /root/reference/data by tests/golden/extract_dynamics_parameters.py)."""
import glob
import json
import os

import numpy as np
import pytest

from conftest import foundation_dr_env_params
from oracle import binding as B

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dynamics_parameters")


@pytest.fixture(scope="module")
def rb():
    import raptor_b200
    return raptor_b200


def golden_files():
    files = sorted(glob.glob(os.path.join(G, "*.json")))
    assert len(files) == 8
    return files


def test_import_golden_dynamics_parameters_matches_plain_json(rb):
    """every number of the reference's files lands in the documented slot of the flat row (checked with Python's own json parser)"""
    for f in golden_files():
        text = open(f).read()
        d = json.loads(text)
        row = rb.parameters_from_json(text)
        dyn = d["dynamics"]
        assert np.array_equal(row[0:12], np.array(dyn["rotor_positions"], np.float32).ravel())
        assert np.array_equal(row[36:48], np.array(dyn["rotor_thrust_coefficients"], np.float32).ravel())
        assert row[60] == np.float32(dyn["mass"]) and np.array_equal(row[64:73], np.array(dyn["J"], np.float32).ravel())
        assert np.array_equal(row[73:82], np.array(dyn["J_inv"], np.float32).ravel())
        assert row[85] == np.float32(d["integration"]["dt"])
        assert row[87] == np.float32(d["mdp"]["init"]["max_position"]) and row[91] == float(d["mdp"]["init"]["relative_rpm"])
        assert row[96] == np.float32(d["mdp"]["reward"]["constant"]) and row[115] == np.float32(d["mdp"]["termination"]["position_threshold"])
        assert row[121] == np.float32(d["disturbances"]["random_force"]["std"])
        assert row[139] == np.float32(d["trajectory"]["mixture"][0]) and row[144] == np.float32(d["trajectory"]["langevin"]["alpha"])
        # export -> import is the identity on the text the reference wrote (6 decimals in, 6 decimals out)
        again = rb.parameters_from_json(rb.parameters_to_json(row))
        assert np.array_equal(again, row)
        back = json.loads(rb.parameters_to_json(row))
        assert back.keys() == d.keys() and back["dynamics"].keys() == dyn.keys() and back["mdp"]["reward"] == d["mdp"]["reward"]


def test_parameters_json_vs_reference(rb, ref):
    if not ref.json_available():
        pytest.skip("reference built without nlohmann/json.hpp")
    rs = np.random.RandomState(1)
    for spec in (B.SPEC_DEFAULT, B.SPEC_RAPTOR, B.SPEC_TEACHER_DR):
        rows = [ref.nominal_parameters(spec)]
        env_p = foundation_dr_env_params(ref, B.SPEC_RAPTOR_DR)
        rng = ref.rng_states(spec, 4, warmup=16)
        rows += list(ref.sample_initial_parameters_n(B.SPEC_RAPTOR_DR, env_p, rng))
        weird = rows[0].copy(); weird[:] = rs.normal(0, 3.0, 145).astype(np.float32); weird[[91, 94, 114]] = [1, 0, 1]
        rows.append(weird)
        for row in rows:
            assert rb.parameters_to_json(row) == ref.parameters_to_json(spec, row)                 # character for character
    base = ref.nominal_parameters(B.SPEC_RAPTOR)
    for f in golden_files():
        text = open(f).read()
        assert np.array_equal(rb.parameters_from_json(text, base), ref.parameters_from_json(B.SPEC_RAPTOR, text, base))


def test_parameters_from_json_errors(rb):
    text = open(golden_files()[0]).read()
    d = json.loads(text)
    row0 = np.full(145, 7.0, np.float32)
    for mutate, needle in [(lambda x: x["dynamics"].pop("mass"), "dynamics.mass"), (lambda x: x["mdp"]["reward"].pop("d_action"), "mdp.reward.d_action"),
                           (lambda x: x.pop("trajectory"), "trajectory"), (lambda x: x["trajectory"].__setitem__("MIXTURE_N", 3), "MIXTURE_N"),
                           (lambda x: x["dynamics"].__setitem__("J", [[1, 2, 3]]), "dynamics.J"), (lambda x: x["mdp"]["init"].__setitem__("relative_rpm", 1.0), "relative_rpm")]:
        bad = json.loads(text); mutate(bad)
        row = row0.copy()
        with pytest.raises(rb.EngineError, match=needle):
            rb.parameters_from_json(json.dumps(bad), row)
        assert np.array_equal(row, row0)
    for broken in ("", "{", text[:-1], text + "x", "[1, 2"):
        with pytest.raises(rb.EngineError):
            rb.parameters_from_json(broken)
    extra = dict(d); extra["unknown"] = {"a": [1, "b", None, True]}
    assert np.array_equal(rb.parameters_from_json(json.dumps(extra)), rb.parameters_from_json(text))


@pytest.mark.gpu
def test_state_json_vs_reference(rb, ref):
    if not ref.json_available():
        pytest.skip("reference built without nlohmann/json.hpp")
    for spec in (B.SPEC_DEFAULT, B.SPEC_RAPTOR):
        env = rb.VectorEnvironment(64, spec)
        env.initialize_rng(9, warmup=16)
        env.sample_initial_state()
        env.load_policy()
        env.rollout(7)
        states, params = env.get_state(), env.get_parameters()
        kinds = set()
        for i in range(64):
            text = env.state_to_json(states[i])
            assert text == ref.state_to_json(spec, params[i], states[i])
            kinds.add(json.loads(text)["trajectory"]["type"])
            base = np.full(env.STATE_DIM, 3.0, np.float32)
            assert np.array_equal(env.state_from_json(text, base), ref.state_from_json(spec, params[i], text, base))
        assert kinds == ({"POSITION"} if spec == B.SPEC_DEFAULT else {"POSITION", "LANGEVIN"})
