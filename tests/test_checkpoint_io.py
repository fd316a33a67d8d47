"""Checkpoint code export (`checkpoint.h`) reader, include/b200_l2f.h b200l2f_checkpoint_*: against files written by the REFERENCE's own
save_code (tests/golden/checkpoints/*.h.gz, made by tests/golden/generate_checkpoints.py through oracle/_ref), against the Raptor checkpoint
file itself where the reference tree is present, and -- on the GPU -- the known-answer pair stored in each file through the engine."""
import gzip
import os

import numpy as np
import pytest

from oracle import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "checkpoints")
RAPTOR_FILE = os.path.join(ROOT, "oracle", "_ref", "ckpt", "checkpoint.h")   # extracted from the reference's tarball by oracle/Makefile (build container only)


def fixture_text(name):
    return gzip.open(os.path.join(FIX, name + ".h.gz")).read()


@pytest.mark.parametrize("name,dims,standardize,head", [("teacher_sac", (26, 64, 8), 0, 1), ("ppo_actor", (22, 64, 4), 1, 2)])
def test_reader_recovers_reference_exports_bit_exactly(name, dims, standardize, head):
    import raptor_b200 as rb
    c = rb.Checkpoint(text=fixture_text(name))
    desc, blob = c.policy()
    want = np.load(os.path.join(FIX, "blobs.npz"))[name]
    assert (desc.arch, desc.input_dim, desc.hidden_dim, desc.output_dim, desc.standardize, desc.head) == (rb.POLICY_MLP,) + dims + (standardize, head)
    assert blob.dtype == np.float32 and np.array_equal(blob.view(np.uint32), want.view(np.uint32))
    assert c.name == "fixtures/" + name and c.commit_hash == "fixture"
    x, y = c.example
    assert x.shape[-1] == dims[0] and y.shape[-1] == 4 and x.shape[:-1] == y.shape[:-1]


@pytest.mark.parametrize("name", ["teacher_sac", "ppo_actor"])
def test_stored_example_is_reproduced_by_the_oracle(name):
    """the blob order the reader assembles is the one the oracle port's MLP consumes: its forward pass of the stored example input gives the
    stored example output (Evaluation mode: tanh(mean) for the SAC actor, the mean for the PPO actor)"""
    import raptor_b200 as rb
    port = B.Port()
    c = rb.Checkpoint(text=fixture_text(name))
    d, blob = c.policy()
    x, y = c.example
    if d.head == rb.HEAD_PPO_GAUSSIAN:
        blob = blob[:-4]                                   # Evaluation mode of the PPO actor: the mean, log_std unused
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=d.input_dim, hidden_dim=d.hidden_dim, output_dim=d.output_dim, standardize=d.standardize,
                           head=B.HEAD_SQUASH_EVAL if d.head == rb.HEAD_SQUASH_EVAL else B.HEAD_IDENTITY)
    got, _, _ = port.policy_evaluate_step(pol, x.reshape(-1, d.input_dim))
    np.testing.assert_allclose(got[:, :4], y.reshape(-1, 4), rtol=2e-6, atol=2e-6)


@pytest.mark.skipif(not os.path.exists(RAPTOR_FILE), reason="reference checkpoint not extracted here")
def test_raptor_checkpoint_file():
    import raptor_b200 as rb
    c = rb.Checkpoint(path=RAPTOR_FILE)
    desc, blob = c.policy()
    assert (desc.arch, desc.input_dim, desc.hidden_dim, desc.output_dim, desc.head, desc.gru_sequence_length) == (rb.POLICY_RAPTOR_GRU, 22, 16, 4, rb.HEAD_IDENTITY, 500)
    g = np.load(os.path.join(ROOT, "tests", "golden", "raptor_kat.npz"))
    assert np.array_equal(blob, rb.raptor_policy_blob()) and np.array_equal(blob, g["blob"])
    x, y = c.example
    assert np.array_equal(x, g["input"]) and np.array_equal(y, g["output"])
    assert c.name == "logs/2025-04-19_16-16-17"
    assert np.array_equal(c.tensors["rl_tools::checkpoint::actor::layer_1::initial_hidden_state"], g["h0"])


def test_reader_errors():
    import raptor_b200 as rb
    with pytest.raises(rb.EngineError, match="no `memory\\[\\]` tensors"):
        rb.Checkpoint(text="namespace a { int x = 3; }")
    with pytest.raises(rb.EngineError, match="unbalanced"):
        rb.Checkpoint(text="namespace a { namespace b { ")
    with pytest.raises(rb.EngineError, match="do not match its shape"):
        rb.Checkpoint(text="namespace a { alignas(float) const unsigned char memory[] = {0, 0, 0, 0}; using SHAPE = x::tensor::Shape<unsigned long, 2>; }")
    with pytest.raises(rb.EngineError, match="byte value"):
        rb.Checkpoint(text="namespace a { alignas(float) const unsigned char memory[] = {0, 0, 0, 300}; using SHAPE = x::tensor::Shape<unsigned long, 1>; }")
    # a well-formed export whose actor is not one of the architectures the engine runs: tensors readable, policy() refuses
    text = fixture_text("teacher_sac").decode().replace("hidden_layer_0", "hidden_layer_9")
    c = rb.Checkpoint(text=text)
    assert "rl_tools::checkpoint::actor::layer_0::hidden_layer_9::weights" in c.tensors
    with pytest.raises(rb.EngineError, match="3-layer MLPs"):
        c.policy()
    # double-precision exports narrow to float; padded row pitch (RowMajorAlignment<TI, 4>) is removed
    import struct
    dbl = ", ".join(str(b) for b in struct.pack("<3d", 1.5, -2.25, 3.0))
    c = rb.Checkpoint(text="namespace n { alignas(double) const unsigned char memory[] = {%s}; using SHAPE = t::Shape<unsigned long, 3>; }" % dbl)
    assert np.array_equal(c.tensors["n"], np.array([1.5, -2.25, 3.0], np.float32))
    pad = ", ".join(str(b) for b in struct.pack("<8f", 1, 2, 3, 99, 4, 5, 6, 99))
    c = rb.Checkpoint(text="namespace m { namespace parameters_memory { alignas(float) const unsigned char memory[] = {%s}; using CONTAINER_SPEC = "
                           "r::matrix::Specification<float, unsigned long, 2, 3, true, r::matrix::layouts::RowMajorAlignment<unsigned long, 4>>; } }" % pad)
    assert np.array_equal(c.tensors["m"], np.array([[1, 2, 3], [4, 5, 6]], np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["teacher_sac", "ppo_actor"])
@pytest.mark.parametrize("gemm", ["tcgen05", "fp32"])
def test_engine_reproduces_stored_example(name, gemm):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200 as rb
    c = rb.Checkpoint(text=fixture_text(name))
    x, y = c.example
    x, y = x.reshape(-1, x.shape[-1]), y.reshape(-1, 4)
    kw = c.policy_kwargs()
    kw["gemm"] = rb.GEMM_TCGEN05_3XTF32 if gemm == "tcgen05" else rb.GEMM_FP32_CUDA_CORES
    if kw["head"] == rb.HEAD_PPO_GAUSSIAN:
        kw["head"], kw["blob"] = rb.HEAD_IDENTITY, kw["blob"][:-4]      # Evaluation mode of the PPO actor: the mean
    env = rb.VectorEnvironment(x.shape[0], rb.SPEC_TEACHER if x.shape[1] == 26 else rb.SPEC_RAPTOR)
    env.load_policy(**kw)
    got = env.policy_evaluate_step(x)
    np.testing.assert_allclose(got, y, rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(RAPTOR_FILE), reason="reference checkpoint not extracted here")
def test_engine_reproduces_raptor_example_from_file():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200 as rb
    c = rb.Checkpoint(path=RAPTOR_FILE)
    x, y = c.example                                      # [500, 2, 22] -> [500, 2, 4]
    env = rb.VectorEnvironment(x.shape[1], rb.SPEC_RAPTOR)
    env.load_policy(**c.policy_kwargs())
    env.policy_reset()
    for t in range(x.shape[0]):
        np.testing.assert_allclose(env.policy_evaluate_step(np.ascontiguousarray(x[t])), y[t], rtol=1e-4, atol=2e-5)
