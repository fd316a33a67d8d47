"""include/rl_tools_b200.h: the `rl_tools::`-namespace binding (device tag rl_tools::devices::B200 + overloads on rl_tools::Matrix / the reference's
own environment, model and evaluation-result types), compiled against the REAL rl-tools headers.

CPU suite (this container, needs /root/reference): tests/cpp/rl_tools_binding.cpp compiles and links -- every rlt:: call in it resolves to the B200
overloads.  The binary is git-ignored but travels to the GPU box, where the -m gpu test runs it: the README loop / rl_tools::evaluate written in the
reference's idiom must reproduce the golden trajectory generated from the reference (tests/golden/default_8x500.npz)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "rl_tools_binding")
REF_INC = "/root/reference/rl-tools/include"
CKPT_INC = os.path.join(ROOT, "oracle", "_ref", "ckpt")


def build_exe():
    from raptor_b200 import build
    from oracle import binding
    build.build()
    binding.build("ref")          # extracts checkpoint.h (the reference's Raptor model as C++ code) into oracle/_ref/ckpt
    cmd = ["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", REF_INC, "-I", CKPT_INC, os.path.join(ROOT, "tests", "cpp", "rl_tools_binding.cpp"),
           "-o", EXE, "-L", os.path.join(ROOT, "raptor_b200", "lib"), "-lb200l2f", "-Wl,-rpath,$ORIGIN/../../raptor_b200/lib"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="needs the reference headers (/root/reference)")
def test_binding_compiles_against_the_reference_headers(tmp_path):
    build_exe()
    assert os.path.exists(EXE)
    # the overloads really are in namespace rl_tools and take the reference's containers: the instantiations of an unoptimised object file say so
    obj = str(tmp_path / "binding.o")
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-c", "-I", os.path.join(ROOT, "include"), "-I", REF_INC, "-I", CKPT_INC, os.path.join(ROOT, "tests", "cpp", "rl_tools_binding.cpp"), "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    syms = subprocess.run(["nm", "-C", obj], capture_output=True, text=True).stdout
    for fn in ("malloc", "init", "sample_initial_parameters", "sample_initial_state", "observe", "step", "evaluate_step", "evaluate", "copy", "reset", "free"):
        assert any(("rl_tools::%s<" % fn) in line and "rl_tools::devices::B200&" in line for line in syms.splitlines()), fn
    assert "rl_tools::Matrix<rl_tools::matrix::Specification<float, unsigned long, 8ul, 82ul" in syms       # the reference's container type in the B200 observe


@pytest.mark.gpu
def test_rl_tools_idiom_loop_matches_golden(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    if not os.path.exists(EXE):
        if not os.path.isdir(REF_INC):
            pytest.skip("binary not built (needs /root/reference at build time)")
        build_exe()
    out = str(tmp_path / "out.bin")
    r = subprocess.run([EXE, out], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out, np.float32)
    T, n = 100, 8
    actions = raw[: T * n * 4].reshape(T, n, 4)
    states = raw[T * n * 4: T * n * 4 + n * 108].reshape(n, 108)
    returns = raw[T * n * 4 + n * 108: T * n * 4 + n * 108 + n]
    stats = raw[T * n * 4 + n * 108 + n:]
    g = np.load(os.path.join(ROOT, "tests", "golden", "default_8x500.npz"))
    scale = np.maximum(np.abs(g["actions"][:T]).max(axis=(0, 2)), 0.1)
    assert (np.abs(actions - g["actions"][:T]).max(axis=(0, 2)) <= 1e-4 * scale).all()
    want = g["states"][list(g["state_steps"]).index(T)]
    for sl, floor in [(slice(0, 3), 0.1), (slice(3, 7), 1.0), (slice(7, 10), 0.1), (slice(10, 13), 0.1), (slice(26, 30), 0.1)]:
        sc = np.maximum(np.abs(want[:, sl]).max(axis=1), floor)
        assert (np.abs(states[:, sl] - want[:, sl]).max(axis=1) <= 1e-4 * sc).all()
    want_returns = g["rewards"][:T].sum(0)
    np.testing.assert_allclose(returns, want_returns, rtol=1e-3, atol=1e-2)
    # the reference's Result aggregates (rl/utils/evaluation/operations_generic.h:201-213)
    np.testing.assert_allclose(stats[0], returns.mean(), rtol=1e-5)
    np.testing.assert_allclose(stats[1], returns.std(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(stats[2], 100.0, rtol=0, atol=1e-6)       # nobody terminates within 100 steps of the hover task
    np.testing.assert_allclose(stats[3], 0.0, rtol=0, atol=1e-3)
    assert stats[4] == 0 and stats[5] == 0                               # num_terminated, share_terminated
