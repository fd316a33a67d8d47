"""Top-level `foundation_policy` package: the import name the reference's README uses (R/README.md:45: `from foundation_policy import Raptor`).
A thin alias of the engine's actor module (pybind11 extension, or the ctypes twin with B200L2F_PYTHON_BINDING=ctypes / when the extension is not built)."""
import importlib
import os

if os.environ.get("B200L2F_PYTHON_BINDING", "pybind") == "ctypes":
    Raptor = importlib.import_module("raptor_b200.foundation_policy").Raptor
else:
    try:
        Raptor = importlib.import_module("raptor_b200._l2f_pybind").foundation_policy.Raptor
    except ImportError:
        Raptor = importlib.import_module("raptor_b200.foundation_policy").Raptor

__all__ = ["Raptor"]
