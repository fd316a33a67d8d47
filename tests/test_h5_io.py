"""`checkpoint.h5` reader (include/b200_l2f.h b200l2f_checkpoint_parse_h5; raptor_b200/csrc/h5_io.cu -- the engine's own HDF5 reader, no libhdf5).

Pinned on the REFERENCE's file: tests/golden/checkpoints/raptor_checkpoint.h5 is the checkpoint.h5 of the published Raptor policy (copied out of
/root/reference/data/raptor-policy-checkpoint.tar.gz by tests/golden/extract_raptor_h5.py; written by rl::loop::steps::checkpoint::save through
HighFive / libhdf5).  Its weights must equal -- bit for bit -- the ones of the code export checkpoint.h next to it in the tarball (committed as
tests/golden/raptor_kat.npz), and the example pair stored in the .h5 (a second known-answer test: `save` draws its own randn input, :140-147)
must be reproduced by the oracle.  Actor shapes and datatypes that file does not cover are written by tests/h5_writer.py (independent of the reader)."""
import gzip
import os

import numpy as np
import pytest

from oracle import binding as B

import h5_writer as W

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "checkpoints")
RAPTOR_H5 = os.path.join(FIX, "raptor_checkpoint.h5")
RAPTOR_H = os.path.join(ROOT, "oracle", "_ref", "ckpt", "checkpoint.h")
P = "rl_tools::checkpoint::"


def raptor_bytes():
    with open(RAPTOR_H5, "rb") as f:
        return f.read()


def test_raptor_h5_gives_the_code_exports_weights_bit_for_bit():
    import raptor_b200 as rb
    c = rb.Checkpoint(path=RAPTOR_H5)
    g = np.load(os.path.join(ROOT, "tests", "golden", "raptor_kat.npz"))
    desc, blob = c.policy()
    assert (desc.arch, desc.input_dim, desc.hidden_dim, desc.output_dim, desc.head) == (rb.POLICY_RAPTOR_GRU, 22, 16, 4, rb.HEAD_IDENTITY)
    assert desc.gru_sequence_length == 0                      # not recorded in the file: the engine's default (500) applies
    assert blob.dtype == np.float32 and np.array_equal(blob.view(np.uint32), g["blob"].view(np.uint32))
    assert np.array_equal(blob, rb.raptor_policy_blob())
    shapes = {k[len(P):]: v.shape for k, v in c.tensors.items()}
    assert shapes == {"actor::layer_0::weights": (16, 22), "actor::layer_0::biases": (1, 16), "actor::layer_1::weights_input": (48, 16),
                      "actor::layer_1::biases_input": (48,), "actor::layer_1::weights_hidden": (48, 16), "actor::layer_1::biases_hidden": (48,),
                      "actor::layer_1::initial_hidden_state": (16,), "actor::layer_2::weights": (4, 16), "actor::layer_2::biases": (1, 4),
                      "example::input": (500, 2, 22), "example::output": (500, 2, 4)}
    assert c.name == "logs/2025-04-19_16-16-17" and c.commit_hash is None


def test_raptor_h5_attributes():
    """what rl_tools::save writes next to the data: layer types, activation functions (dense/persist.h:17-18), the container attributes
    (matrix/persist.h:21-23, tensor/persist.h:88-100) and the environment description (`meta`, operations_cpu.h:128-131)"""
    import json
    import raptor_b200 as rb
    c = rb.Checkpoint(path=RAPTOR_H5)
    s = c.string
    assert s(P + "actor::type") == "sequential"
    assert [s(P + "actor::layer_%d::type" % k) for k in range(3)] == ["dense", "gru", "dense"]
    assert s(P + "actor::layer_0::activation_function") == "RELU" and s(P + "actor::layer_2::activation_function") == "IDENTITY"
    assert (s(P + "actor::layer_0::weights::type"), s(P + "actor::layer_0::weights::rows"), s(P + "actor::layer_0::weights::cols")) == ("matrix", "16", "22")
    assert (s(P + "actor::layer_1::weights_input::type"), s(P + "actor::layer_1::weights_input::num_dims"), s(P + "actor::layer_1::weights_input::dim_0")) == ("tensor", "2", "48")
    meta = json.loads(s(P + "actor::meta"))
    assert meta["environment"]["name"] == "l2f" and meta["environment"]["observation"].startswith("Position.Orientation")
    assert s(P + "actor::nothing") is None


def test_raptor_h5_example_is_a_known_answer(port):
    """the pair `save` stores in the .h5 is NOT the one `save_code` stores in checkpoint.h (each draws its own input): a second KAT from the reference"""
    import raptor_b200 as rb
    c = rb.Checkpoint(path=RAPTOR_H5)
    x, y = c.example
    kat = np.load(os.path.join(ROOT, "tests", "golden", "raptor_kat.npz"))
    assert x.shape == kat["input"].shape and not np.array_equal(x, kat["input"])
    assert abs(float(x.mean())) < 0.02 and abs(float(x.std()) - 1.0) < 0.02          # randn (operations_cpu.h:143)
    desc, blob = c.policy()
    pol = port.make_policy(blob)
    h0 = c.tensors[P + "actor::layer_1::initial_hidden_state"]
    errs = []
    for b in range(2):
        h, st = h0[None].astype(np.float32).copy(), np.zeros(1, np.int32)
        for t in range(500):
            a, _, _ = port.policy_evaluate_step(pol, x[t, b:b + 1], h, st)
            errs.append(np.abs(a[0] - y[t, b]))
    errs = np.array(errs)
    assert errs.mean() < 5e-7 and errs.max() < 3e-6


@pytest.mark.skipif(not os.path.exists(RAPTOR_H), reason="reference checkpoint not extracted here")
def test_raptor_h5_against_the_code_export_file():
    import raptor_b200 as rb
    a, b = rb.Checkpoint(path=RAPTOR_H), rb.Checkpoint(path=RAPTOR_H5)
    for k, v in a.tensors.items():
        if "::example::" in k:
            continue
        assert b.tensors[k].shape == v.shape and np.array_equal(b.tensors[k].view(np.uint32), v.view(np.uint32)), k
    assert a.name == b.name


# ---- actor shapes the reference file does not cover: the code-export fixtures re-laid out as rl_tools::save would write them ----------------------
def h5_tree_from_export(c, kind):
    """tensors of a parsed code export -> the group tree of checkpoint.h5 (nn_models/sequential/persist.h:14-21, nn_models/mlp/persist.h:14-22,
    nn_models/mlp_unconditional_stddev/persist.h:12-15, nn/layers/{dense,standardize,sample_and_squash}/persist.h, nn/parameters/persist.h:10-13)"""
    root = W.Group()
    for path, arr in c.tensors.items():
        parts = path[len(P):].split("::")
        node = root
        comps = []
        for p in parts:
            if p.startswith("layer_") and p[6:].isdigit():
                comps += ["layers", p[6:]]
            else:
                comps.append(p)
        leaf_is_parameter = comps[0] == "actor"
        for p in comps[:-1] if not leaf_is_parameter else comps:
            node = node.children.setdefault(p, W.Group())
        attrs = {"type": "matrix", "rows": str(arr.shape[0]), "cols": str(arr.shape[1])} if arr.ndim == 2 else \
                dict({"type": "tensor", "num_dims": str(arr.ndim)}, **{"dim_%d" % i: str(d) for i, d in enumerate(arr.shape)})
        node.children["parameters" if leaf_is_parameter else comps[-1]] = W.Dataset(arr, attrs)
    actor = root.children["actor"]
    actor.attrs.update(type="sequential", checkpoint_name=c.name, meta='{"environment": {"name": "l2f"}}')
    layers = actor.children["layers"].children
    mlp = layers["0" if kind == "teacher_sac" else "1"]
    mlp.attrs.update(type="mlp", num_layers="3")
    for name, act in (("input_layer", "RELU"), ("hidden_layer_0", "RELU"), ("output_layer", "IDENTITY")):
        mlp.children[name].attrs.update(activation_function=act, type="dense")
    if kind == "teacher_sac":
        layers["1"] = W.Group(attrs={"type": "sample_and_squash"})
    return root


@pytest.mark.parametrize("fixed_strings", [False, True])
@pytest.mark.parametrize("name,dims,standardize,head", [("teacher_sac", (26, 64, 8), 0, 1), ("ppo_actor", (22, 64, 4), 1, 2)])
def test_mlp_actors_in_the_h5_layout(name, dims, standardize, head, fixed_strings):
    import raptor_b200 as rb
    src = rb.Checkpoint(text=gzip.open(os.path.join(FIX, name + ".h.gz")).read())
    data = W.write(h5_tree_from_export(src, name), fixed_strings=fixed_strings)
    c = rb.Checkpoint(text=data)
    desc, blob = c.policy()
    want = np.load(os.path.join(FIX, "blobs.npz"))[name]
    assert (desc.arch, desc.input_dim, desc.hidden_dim, desc.output_dim, desc.standardize, desc.head) == (rb.POLICY_MLP,) + dims + (standardize, head)
    assert np.array_equal(blob.view(np.uint32), want.view(np.uint32))
    assert c.name == "fixtures/" + name
    for k, v in src.tensors.items():
        assert np.array_equal(c.tensors[k], v) and c.tensors[k].shape == v.shape, k


def test_datatypes_layouts_and_wide_groups():
    """float64 and big-endian data are converted, integers too; compact layout; a group with more links than one symbol-table node holds;
    empty datasets; a file behind a user block (base address != 0)"""
    import raptor_b200 as rb
    rs = np.random.RandomState(3)
    a = rs.standard_normal((5, 7)).astype(np.float32)
    tree = {"example": W.Group({
        "f32": a, "f64": W.Dataset(a.astype(np.float64), dtype="<f8"), "f32_be": W.Dataset(a, dtype=">f4"), "f64_be": W.Dataset(a.astype(np.float64), dtype=">f8"),
        "compact": W.Dataset(a, compact=True), "i32": W.Dataset(np.arange(-6, 6, dtype=np.int32).reshape(3, 4)), "u8": W.Dataset(np.arange(250, 256, dtype=np.uint8)),
        "i16_be": W.Dataset(np.array([-2, 300], np.int16), dtype=">i2"), "empty": W.Dataset(np.zeros((0, 4), np.float32)), "scalar": W.Dataset(np.float32(2.5)),
        "wide": W.Group({"d%02d" % i: np.full((2,), i, np.float32) for i in range(37)}),
    })}
    for userblock, version in ((0, 0), (512, 0), (2048, 1), (0, 1)):
        c = rb.Checkpoint(text=W.write(tree, userblock=userblock, superblock_version=version))
        t = {k[len(P + "example::"):]: v for k, v in c.tensors.items()}
        for k in ("f32", "f64", "f32_be", "f64_be", "compact"):
            assert np.array_equal(t[k].view(np.uint32), a.view(np.uint32)), k
        assert np.array_equal(t["i32"], np.arange(-6, 6, dtype=np.float32).reshape(3, 4)) and np.array_equal(t["u8"], np.arange(250, 256, dtype=np.float32))
        assert np.array_equal(t["i16_be"], np.array([-2, 300], np.float32))
        assert t["empty"].shape == (0, 4) and t["scalar"].shape == () and float(t["scalar"]) == 2.5
        assert sorted(k for k in t if k.startswith("wide::")) == ["wide::d%02d" % i for i in range(37)]
        assert all(np.array_equal(t["wide::d%02d" % i], np.full((2,), i, np.float32)) for i in range(37))
        with pytest.raises(rb.EngineError, match="no actor|not of a kind|unsupported"):
            c.policy()


def test_what_the_reader_does_not_read_is_named():
    import raptor_b200 as rb
    good = raptor_bytes()
    with pytest.raises(rb.EngineError, match="superblock version 2"):
        rb.Checkpoint(text=good[:8] + b"\x02" + good[9:])
    with pytest.raises(rb.EngineError, match="size of offsets / lengths 4 / 8"):
        rb.Checkpoint(text=good[:13] + b"\x04" + good[14:])
    chunked = bytearray(W.write({"example": W.Group({"x": np.ones((4,), np.float32)})}))
    import struct
    at = chunked.find(struct.pack("<HHB3x", 8, 24, 0) + bytes([3, 1])) + 8                   # the layout message: version 3, class 1 (contiguous)
    assert at > 8
    chunked[at + 1] = 2
    with pytest.raises(rb.EngineError, match="chunked dataset '/example/x'"):
        rb.Checkpoint(text=bytes(chunked))
    with pytest.raises(rb.EngineError, match="no `memory\\[\\]` tensors|not an rl-tools"):
        rb.Checkpoint(text=b"\x89HDF but not really")
    lib = rb._lib.load() if hasattr(rb, "_lib") else None
    if lib is not None:
        import ctypes
        h = ctypes.c_void_p()
        assert lib.b200l2f_checkpoint_parse_h5(b"\x89HDF but not really", 19, ctypes.byref(h)) != 0 and not h.value
        assert b"not an HDF5 file" in lib.b200l2f_last_error(None)


def test_truncated_and_corrupted_files_fail_cleanly():
    """every access of the reader is bounds-checked: cutting the reference's file anywhere, or flipping bytes in its metadata, gives an error or a
    (possibly different) result -- never a crash.  Runs in a child process so that a crash would be seen as one."""
    import subprocess
    import sys
    code = r'''
import ctypes, sys, numpy as np
sys.path.insert(0, %r)
from raptor_b200 import _lib
lib = _lib.load()
good = open(%r, "rb").read()
def attempt(b):
    h = ctypes.c_void_p()
    rc = lib.b200l2f_checkpoint_parse_h5(b, len(b), ctypes.byref(h))
    if rc == 0:
        n = lib.b200l2f_checkpoint_tensor_count(h); lib.b200l2f_checkpoint_free(h)
    return rc
assert attempt(good) == 0
failed = 0
for cut in list(range(0, 4096, 7)) + list(range(4096, len(good), 997)):
    failed += attempt(good[:cut]) != 0
rs = np.random.RandomState(0)
meta = 36000                                     # everything before the example data is object headers / trees / heaps / small datasets
for trial in range(600):
    b = bytearray(good)
    for _ in range(rs.randint(1, 4)):
        b[rs.randint(8, meta)] = rs.randint(0, 256)
    failed += attempt(bytes(b)) != 0
# global heap objects whose 64-bit size field is near 2^64 (once wrapped the bounds check into std::length_error / an endless scan)
g = good.find(b"GCOL")
assert g > 0
for obj in range(3):
    for size in (0xFFFFFFFFFFFFFFF0, 0xFFFFFFFFFFFFFFF8, 0xFFFFFFFFFFFFFFFF, 0x8000000000000000):
        b = bytearray(good)
        at = g + 16
        for _ in range(obj):                                     # walk to object `obj` of the first collection
            at += 16 + ((int.from_bytes(good[at + 8:at + 16], "little") + 7) & ~7)
        b[at + 8:at + 16] = size.to_bytes(8, "little")
        assert attempt(bytes(b)) != 0, (obj, hex(size))
print("ok", failed)
''' % (ROOT, RAPTOR_H5)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
    assert int(r.stdout.split()[1]) > 300                       # most mutilated files are rejected (the rest hit padding or unread fields)


def test_reader_under_address_sanitizer(tmp_path):
    """the reader compiled as plain C++ with -fsanitize=address,undefined, fed ~13 000 truncated / byte-flipped variants of the reference's file
    (tests/cpp/h5_fuzz.cpp): no out-of-range access, no undefined behaviour"""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "h5_fuzz")
    csrc = os.path.join(ROOT, "raptor_b200", "csrc")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", csrc, "-x", "c++",
                        os.path.join(csrc, "h5_io.cu"), os.path.join(ROOT, "tests", "cpp", "h5_fuzz.cpp"), "-o", exe], capture_output=True, text=True)
    if r.returncode != 0 and "sanitize" in r.stderr + r.stdout:
        pytest.skip("this g++ has no sanitizer runtime")
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, RAPTOR_H5, "36000", "5000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


@pytest.mark.gpu
def test_engine_reproduces_the_h5_example_on_gpu():
    """the .h5's own known-answer pair through the CUDA engine (actor loaded from the .h5, both via VectorEnvironment and foundation_policy.Raptor)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200 as rb
    from raptor_b200 import foundation_policy
    c = rb.Checkpoint(path=RAPTOR_H5)
    x, y = c.example                                      # [500, 2, 22] -> [500, 2, 4]
    env = rb.VectorEnvironment(x.shape[1], rb.SPEC_RAPTOR)
    env.load_policy(**c.policy_kwargs())
    env.policy_reset()
    policy = foundation_policy.Raptor(checkpoint=RAPTOR_H5)
    policy.reset()
    for t in range(x.shape[0]):
        obs = np.ascontiguousarray(x[t])
        np.testing.assert_allclose(env.policy_evaluate_step(obs), y[t], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(policy.evaluate_step(obs), y[t], rtol=1e-4, atol=2e-5)
