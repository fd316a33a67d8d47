"""Header-only C++ shim (include/b200_l2f.hpp, rl-tools idiom) : compiles everywhere (no GPU needed to build), and on the GPU box runs the
README loop in C++ and matches the golden trajectory of BASELINE config 1."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "readme_loop")
PPO_EXE = os.path.join(ROOT, "tests", "cpp", "ppo_loop")


def build_exe(name="readme_loop"):
    from raptor_b200 import build
    build.build()
    lib_dir = os.path.join(ROOT, "raptor_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
           "-o", os.path.join(ROOT, "tests", "cpp", name), "-L", lib_dir, "-lb200l2f", "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True)


def test_header_shim_compiles_and_links():
    build_exe()
    assert os.path.exists(EXE)


def test_training_side_shim_compiles_and_reads_checkpoints(tmp_path):
    """the runner / learner-feed / DAgger / checkpoint / JSON overloads of the shim compile (-Wall -Werror); its host-only part -- b200::load of a
    checkpoint file, either format -- runs without a GPU and gives the published Raptor actor"""
    import raptor_b200 as rb
    build_exe("ppo_loop")
    r = subprocess.run([PPO_EXE, "checkpoint", os.path.join(ROOT, "tests", "golden", "checkpoints", "raptor_checkpoint.h5")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    want = "arch %d in 22 hidden 16 out 4 floats 2084 sum %.9g name logs/2025-04-19_16-16-17" % (rb.POLICY_RAPTOR_GRU, float(np.sum(rb.raptor_policy_blob().astype(np.float64))))
    assert r.stdout.strip() == want, (r.stdout, want)
    bad = tmp_path / "not_a_checkpoint.h"
    bad.write_text("namespace a { int x = 3; }")
    r = subprocess.run([PPO_EXE, "checkpoint", str(bad)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "memory[]" in r.stderr          # rl-tools idiom: errors terminate through assert_exit
    # b200::json / from_json on a parameter row: the text is the engine's (= the reference's, tests/test_json_io.py)
    from oracle import binding as B
    row = B.Port().nominal_parameters(B.SPEC_RAPTOR).astype(np.float32)
    (tmp_path / "row.f32").write_bytes(row.tobytes())
    r = subprocess.run([PPO_EXE, "json", str(tmp_path / "row.f32"), str(tmp_path / "row.json"), str(tmp_path / "again.f32")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    text = (tmp_path / "row.json").read_text()
    assert text == rb.parameters_to_json(row) and r.stdout.strip() == "%d characters" % len(text)
    again = np.fromfile(str(tmp_path / "again.f32"), np.float32)
    assert np.array_equal(again, rb.parameters_from_json(text, np.full(145, -7.0, np.float32)))
    np.testing.assert_allclose(again, row, rtol=0, atol=6e-7)             # std::to_string keeps 6 decimals (L2F/operations_cpu.h:139-411)


@pytest.mark.gpu
def test_cpp_readme_loop_matches_golden(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if not os.path.exists(EXE):
        build_exe()
    out = str(tmp_path / "out.bin")
    r = subprocess.run([EXE, os.path.join(ROOT, "raptor_b200", "data", "raptor_policy_2084.f32"), out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, np.float32)
    T, n = 100, 8
    actions = raw[: T * n * 4].reshape(T, n, 4)
    states = raw[T * n * 4: T * n * 4 + n * 108].reshape(n, 108)
    returns = raw[T * n * 4 + n * 108:]
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_8x500.npz"))
    scale = np.maximum(np.abs(g["actions"][:T]).max(axis=(0, 2)), 0.1)
    assert (np.abs(actions - g["actions"][:T]).max(axis=(0, 2)) <= 1e-4 * scale).all()
    want = g["states"][list(g["state_steps"]).index(T)]
    for sl, floor in [(slice(0, 3), 0.1), (slice(3, 7), 1.0), (slice(7, 10), 0.1), (slice(10, 13), 0.1), (slice(26, 30), 0.1)]:
        sc = np.maximum(np.abs(want[:, sl]).max(axis=1), floor)
        assert (np.abs(states[:, sl] - want[:, sl]).max(axis=1) <= 1e-4 * sc).all()
    np.testing.assert_allclose(returns, g["rewards"][:T].sum(0), rtol=1e-3, atol=1e-2)
