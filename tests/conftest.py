"""pytest configuration: markers + shared fixtures (oracle handles are TEST infrastructure)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import binding
    binding.build("port")
    return binding.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import binding
    if os.path.isdir("/root/reference/rl-tools/include"):
        binding.build("ref")
    if not binding.Ref.available():
        pytest.skip("oracle/_ref/libl2f_ref.so not built (needs /root/reference)")
    return binding.Ref()


# DR ranges used to train the Raptor teachers (src/foundation_policy/pre_training/sample_dynamics_parameters.cpp:48-64)
FOUNDATION_DR = dict(t2w=(1.5, 5.0), t2i=(40, 1200), mass=(0.02, 5.0), size_dev=0.1, tau_rise=(0.03, 0.10), tau_fall=(0.03, 0.30), kq=(0.005, 0.05), dist_force=0.3)


def foundation_dr_env_params(lib, spec):
    """nominal parameters of `spec` with the foundation-policy DR ranges and its reward constant filled in"""
    p = lib.nominal_parameters(spec).copy()
    d = FOUNDATION_DR
    p[124:139] = np.array([d["t2w"][0], d["t2w"][1], d["t2i"][0], d["t2i"][1], d["mass"][0], d["mass"][1], d["size_dev"],
                           d["tau_rise"][0], d["tau_rise"][1], d["tau_fall"][0], d["tau_fall"][1], d["kq"][0], d["kq"][1], 0.0, d["dist_force"]], np.float32)
    return p


def random_mlp_blob(rs, in_dim, out_dim, standardize, log_std):
    hd = 64
    parts = []
    if standardize:
        parts += [rs.normal(0, 0.1, in_dim), 1.0 / rs.uniform(0.5, 2.0, in_dim)]
    for (o, i) in [(hd, in_dim), (hd, hd), (out_dim, hd)]:
        bound = np.sqrt(6.0 / i)
        w = rs.uniform(-bound, bound, (o, i))
        if o == out_dim:
            w *= 0.3
        parts += [w.ravel(), rs.uniform(-0.05, 0.05, o)]
    if log_std:
        parts.append(np.log(np.full(4, 0.5)))
    return np.concatenate(parts).astype(np.float32)
