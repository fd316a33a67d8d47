"""Top-level `l2f` package: the import name the reference's README uses (R/README.md:43-44: `import l2f`, `from l2f import vector8 as vector`).

A thin alias of the engine's README-compatible module -- the pybind11 extension `raptor_b200._l2f_pybind` (host C++ over the C ABI, the same layering
as the pip wheel) or, with B200L2F_PYTHON_BINDING=ctypes or when the extension has not been built, its ctypes twin `raptor_b200.l2f`.  Both run
every call on the CUDA engine; there is no CPU implementation behind this name."""
import importlib
import os
import sys

if os.environ.get("B200L2F_PYTHON_BINDING", "pybind") == "ctypes":
    _impl = importlib.import_module("raptor_b200.l2f")
else:
    try:
        _impl = importlib.import_module("raptor_b200._l2f_pybind")
    except ImportError:
        _impl = importlib.import_module("raptor_b200.l2f")

Device = _impl.Device
UI = _impl.UI


def vector(n):
    """l2f.vector(N): the vectorN module for any N"""
    mod = _impl.vector(int(n))
    sys.modules.setdefault("l2f.vector%d" % int(n), mod)
    return mod


def __getattr__(name):   # l2f.vector8, l2f.vector64, ...
    if name.startswith("vector") and name[6:].isdigit():
        return vector(int(name[6:]))
    raise AttributeError("module 'l2f' has no attribute %r" % name)
