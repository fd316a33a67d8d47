"""Header-only C++ shim (include/b200_l2f.hpp, rl-tools idiom) : compiles everywhere (no GPU needed to build), and on the GPU box runs the
README loop in C++ and matches the golden trajectory of BASELINE config 1."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "readme_loop")


def build_exe():
    from raptor_b200 import build
    build.build()
    lib_dir = os.path.join(ROOT, "raptor_b200", "lib")
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "readme_loop.cpp"),
           "-o", EXE, "-L", lib_dir, "-lb200l2f", "-Wl,-rpath," + lib_dir]
    subprocess.run(cmd, check=True)


def test_header_shim_compiles_and_links():
    build_exe()
    assert os.path.exists(EXE)


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
