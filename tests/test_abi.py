"""CPU-only checks of the drop-in boundary: the shared library loads and exports every symbol include/b200_l2f.h declares,
the ctypes table matches the header, and the product package never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200_l2f.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200l2f_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from raptor_b200 import build, _lib
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SYMBOLS) == names


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import raptor_b200
    with pytest.raises(raptor_b200.EngineError) as e:
        raptor_b200.VectorEnvironment(8)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, "raptor_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "l2f_oracle" not in txt and "libl2f_ref" not in txt, f


def test_policy_blob_matches_golden():
    import numpy as np
    import raptor_b200
    blob = raptor_b200.raptor_policy_blob()
    g = np.load(os.path.join(ROOT, "tests", "golden", "raptor_kat.npz"))
    assert blob.shape == (2084,) and np.array_equal(blob, g["blob"])


def test_pybind_module_surface():
    """the compiled l2f / foundation_policy modules (csrc/pybind_l2f.cpp) build, import without a GPU and expose the README's names"""
    from raptor_b200 import build
    build.build()
    build.build_pybind()
    import raptor_b200._l2f_pybind as l2f
    vector = l2f.vector8
    for name in ("VectorRng", "VectorEnvironment", "VectorParameters", "VectorState", "initialize_rng", "initialize_environment", "sample_initial_parameters",
                 "sample_initial_state", "observe", "step", "set_ui_message", "set_parameters_message", "set_state_action_message"):
        assert hasattr(vector, name), name
    assert vector.N_ENVIRONMENTS == 8 and l2f.vector(24).N_ENVIRONMENTS == 24 and l2f.vector(24) is l2f.vector24
    assert hasattr(l2f, "Device") and hasattr(l2f, "UI") and hasattr(l2f.foundation_policy, "Raptor")
    ui = l2f.UI(); ui.ns = "abc"
    assert ui.ns == "abc"
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            vector.VectorEnvironment()
