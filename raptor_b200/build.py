"""Builds the engine's shared library in-tree with nvcc for sm_100a (no torch involved in the .so).

    python -m raptor_b200.build            # raptor_b200/lib/libb200l2f.so

The engine is several translation units (one per family of heavy kernel instantiations) compiled in parallel and linked into ONE
shared library; only the units whose sources changed are recompiled.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
LIB = os.environ.get("B200L2F_BUILD_LIB") or os.path.join(LIB_DIR, "libb200l2f.so")
COMMON = ["layout.h", "fastmath.cuh", "rng.cuh", "env.cuh", "samplers.cuh", "policy.cuh", "kernels.cuh", "handle.h", "launch.h", "mlp.cuh",
          os.path.join("..", "..", "include", "b200_l2f.h")]
TC = ["tc.cuh", "rollout_tc.cuh"]
# translation unit -> the headers it depends on besides COMMON
SOURCES = {
    "engine.cu": TC,
    "rollout_fp32.cu": [],
    "rollout_tc.cu": TC,
    "rollout_ts.cu": TC,
    "rollout_x2.cu": TC + ["rollout_x2.cuh"],
    "rollout_mlp.cu": [],
    "rollout_mlp_ts.cu": TC + ["mlp_tc.cuh"],
    "collect.cu": [],
    "collect_ts.cu": TC + ["mlp_tc.cuh", "collect_lag.cuh"],
    "collect_ts_default.cu": TC + ["mlp_tc.cuh"],
    "learner.cu": TC + ["mlp_tc.cuh", "learner.cuh"],
    "dagger.cu": TC + ["mlp_tc.cuh", "dagger.cuh"],
    "off_policy.cu": TC + ["mlp_tc.cuh", "offpolicy.cuh", "offpolicy_tc.cuh"],
    "step_only.cu": TC,
    "collective.cu": [],
    "json_io.cu": [],
    "checkpoint_io.cu": ["h5_io.h"],
    "h5_io.cu": ["h5_io.h"],
}
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC"]
# per-unit flags.  collect_ts_default.cu: every global load through L2 (ld.cg) -- its kernel hands a tile over between SMs inside one launch (time-chunked collection)
UNIT_FLAGS = {"collect_ts_default.cu": ["-Xptxas", "-dlcm=cg"]}


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the engine is CUDA-only and cannot be built without the CUDA toolkit")
    return exe


def _obj(src):
    return os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _unit_stale(src):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, d) for d in [src] + COMMON + SOURCES[src]]
    return any(_mtime(d) > t for d in deps)


def stale():
    return not os.path.exists(LIB) or any(_unit_stale(s) or _mtime(_obj(s)) > os.path.getmtime(LIB) for s in SOURCES)


def build(force=False, verbose=False, extra_flags=(), jobs=None):
    if not force and not stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    todo = [s for s in SOURCES if force or extra_flags or _unit_stale(s)]

    def compile_unit(src):
        cmd = [nvcc()] + NVCC_FLAGS + UNIT_FLAGS.get(src, []) + list(extra_flags) + ["-c", "-o", _obj(src), os.path.join(CSRC, src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs or min(len(todo) or 1, os.cpu_count() or 1)) as pool:
        for src, r in pool.map(compile_unit, todo):
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(s) for s in SOURCES] + ["-ldl"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return LIB


def pybind_path():
    import sysconfig
    return os.path.join(HERE, "_l2f_pybind" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pybind(force=False, verbose=False):
    """the pybind11 twin of raptor_b200/l2f.py + foundation_policy.py (csrc/pybind_l2f.cpp, host C++ over the C ABI): g++ + pybind11 headers,
    linked against the engine library next to it.  Returns the path of the extension module."""
    import sysconfig
    import pybind11
    out = pybind_path()
    src = os.path.join(CSRC, "pybind_l2f.cpp")
    deps = [src, os.path.join(HERE, "..", "include", "b200_l2f.h")]
    if not force and os.path.exists(out) and all(_mtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cxx = shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], src, "-o", out,
           "-L" + LIB_DIR, "-lb200l2f", "-Wl,-rpath,$ORIGIN/lib",
           # keep this module's C++ runtime symbols bound inside it (a statically linked libstdc++ must not be interposed by the process's own)
           "-Wl,-Bsymbolic", "-Wl,--exclude-libs,ALL"]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed on pybind_l2f.cpp")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra_flags=["-Xptxas", "-v"] if "--ptxas" in sys.argv else [])
    print(LIB)
    print(build_pybind(force="--force" in sys.argv, verbose=True))
