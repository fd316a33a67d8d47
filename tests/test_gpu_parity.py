"""GPU parity tests proper (-m gpu): the CUDA engine, called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerances: integer work (RNG streams, ring-buffer indices, GRU step counters, terminated flags) is bit-exact; float32 state/action
trajectories are compared with rtol 1e-4 (the north-star's bound over 100 closed-loop steps) plus a small absolute floor for values
that cross zero; single steps from identical inputs are held to 2e-6 relative."""
import os

import numpy as np
import pytest

from conftest import foundation_dr_env_params, random_mlp_blob
from oracle import binding as B

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rb():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import raptor_b200
    return raptor_b200


def close(a, b, rtol, atol, what=""):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)


# field groups of a state row that share a physical unit: |error| is judged relative to the group's magnitude along the trajectory
STATE_GROUPS = {"position": (slice(0, 3), 0.1), "orientation": (slice(3, 7), 1.0), "linear_velocity": (slice(7, 10), 0.1),
                "angular_velocity": (slice(10, 13), 0.1), "last_action": (slice(13, 17), 0.1), "rpm": (slice(26, 30), 0.1)}


def close_relative(got, want, rtol, groups, what=""):
    """got/want: [T, n, D].  For every env and field group g: max_t |got - want| <= rtol * max(max_t |want_g|, floor_g),
    i.e. "within rtol relative fp32" with the magnitude of the physical quantity as the reference scale."""
    for name, (sl, floor) in groups.items():
        g, w = got[..., sl], want[..., sl]
        scale = np.maximum(np.abs(w).max(axis=(0, 2)), floor)            # [n]
        err = np.abs(g - w).max(axis=(0, 2))                             # [n]
        bad = err > rtol * scale
        assert not bad.any(), "%s/%s: %d envs exceed rtol %g (worst err/scale = %.3g)" % (what, name, bad.sum(), rtol, (err / scale).max())


def test_rng_streams_bit_exact(rb, port):
    env = rb.VectorEnvironment(1000, rb.SPEC_RAPTOR, first_env_id=12345)
    env.initialize_rng(seed=99, warmup=7)
    want = port.rng_states(99, 1000, first_env=12345, warmup=0)
    for i in range(1000):   # warmup advances the raw engine (uniform draws advance it once each)
        s = want[i:i + 1].copy()
        for _ in range(7):
            port.rng_uniform(s, 0, 1)
        want[i] = s[0]
    assert np.array_equal(env.get_rng(), want)


@pytest.mark.parametrize("spec", [B.SPEC_DEFAULT_DR, B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR])
def test_samplers(rb, port, spec):
    n = 512
    env = rb.VectorEnvironment(n, spec)
    env_p = foundation_dr_env_params(port, spec)
    env.set_environment_parameters(env_p)
    env.initialize_rng(seed=3, warmup=16)
    rng = env.get_rng()
    env.sample_initial_parameters()
    want_p = port.sample_initial_parameters_n(spec, env_p, rng)
    got_p = env.get_parameters()
    assert np.array_equal(env.get_rng(), rng)
    close(got_p, want_p, 2e-6, 0, "sampled parameters")
    env.set_parameters(want_p)     # continue from identical parameters
    env.sample_initial_state()
    want_s = port.sample_initial_state_n(spec, want_p, rng)
    assert np.array_equal(env.get_rng(), rng)
    close(env.get_state(), want_s, 2e-6, 1e-7, "sampled states")
    env.initial_state(slot=1)
    want_i = np.array([port.initial_state(spec, want_p[i]) for i in range(n)])
    close(env.get_state(slot=1), want_i, 1e-6, 0, "initial states")


@pytest.mark.parametrize("spec", [B.SPEC_DEFAULT, B.SPEC_RAPTOR, B.SPEC_TEACHER])
@pytest.mark.parametrize("noise", [False, True])
def test_vector_api_single_steps(rb, port, spec, noise):
    """observe / step / reward / terminated from identical inputs, re-anchored on the oracle every step"""
    n, T = 256, 24
    rs = np.random.RandomState(spec)
    env = rb.VectorEnvironment(n, spec)
    p = port.nominal_parameters(spec)
    if noise:
        p[108:113] = [0.01, 0.02, 0.03, 0.04, 0.05]
        p[113] = 0.05
    params = np.tile(p, (n, 1))
    rng = port.rng_states(5, n, warmup=24)
    s = port.sample_initial_state_n(spec, params, rng)
    env.set_parameters(params)
    for t in range(T):
        env.set_state(s)
        env.set_rng(rng)
        a = rs.uniform(-1.2, 1.2, (n, 4)).astype(np.float32)
        obs = env.observe()
        dts = env.step(a)
        nxt = env.get_state(slot=1)
        rew = env.reward(a)
        term = env.terminated(slot=1)
        got_rng = env.get_rng()
        w_obs = np.zeros_like(obs); w_nxt = np.zeros_like(nxt); w_rew = np.zeros(n, np.float32); w_term = np.zeros(n, np.uint8)
        for i in range(n):
            r = rng[i:i + 1].copy()
            w_obs[i] = port.observe(spec, p, s[i], r)
            w_nxt[i], dt = port.step(spec, p, s[i], a[i], r)
            w_rew[i] = port.reward(spec, p, s[i], a[i], w_nxt[i])
            w_term[i] = port.terminated(spec, p, w_nxt[i])
            rng[i] = r[0]
            assert dts[i] == dt
        assert np.array_equal(got_rng, rng)
        tol = 2e-5 if noise else 2e-6   # logf/cosf of the Box-Muller transform differ by an ulp between libm and CUDA
        close(obs, w_obs, tol, tol, "observe")
        close(nxt, w_nxt, 5e-6 if not noise else 5e-5, 2e-6 if not noise else 2e-5, "step")
        close(rew, w_rew, 1e-5, 2e-5, "reward")
        assert np.array_equal(term, w_term)
        s = w_nxt


def test_policy_known_answer_test(rb):
    """the KAT that ships in the checkpoint, through the GPU actor (2 sequences x 500 steps)"""
    k = np.load(os.path.join(G, "raptor_kat.npz"))
    for flags, bound in [(rb.FLAG_ACCURATE_MATH, 3e-6), (0, 1e-5)]:
        env = rb.VectorEnvironment(2, rb.SPEC_RAPTOR, flags=flags)
        env.load_policy(k["blob"])
        env.policy_reset()
        errs = []
        for t in range(500):
            a = env.policy_evaluate_step(np.ascontiguousarray(k["input"][t]))
            errs.append(np.abs(a - k["output"][t]))
        errs = np.array(errs)
        assert errs.max() < bound, (flags, errs.max())
        h, g = env.get_hidden()
        assert np.all(g == 0)   # 500 steps == SEQUENCE_LENGTH: the counter wrapped and the hidden state was reset (gru/operations_generic.h:400-410)
        assert np.array_equal(h, np.tile(k["h0"], (2, 1)))


@pytest.mark.parametrize("name,T_cmp", [("default_8x500.npz", 100), ("raptor_dr_64x100.npz", 100), ("raptor_noise_8x50.npz", 50)])
@pytest.mark.parametrize("flags", [0, 1])
def test_fused_rollout_vs_golden(rb, name, T_cmp, flags):
    """closed loop from identical seeds: states and actions within 1e-4 relative over 100 steps, RNG / counters bit-exact"""
    g = np.load(os.path.join(G, name))
    spec = int(g["spec"])
    n = g["states0"].shape[0]
    T = g["actions"].shape[0]
    env = rb.VectorEnvironment(n, spec, flags=flags)
    env.set_parameters(np.ascontiguousarray(g["params"]))
    env.set_state(np.ascontiguousarray(g["states0"]))
    env.set_rng(np.ascontiguousarray(g["rng0"]))
    env.load_policy(np.load(os.path.join(G, "raptor_kat.npz"))["blob"])
    out = env.rollout(T, record=("states", "actions", "rewards", "terminated"))
    noise = "noise" in name
    rtol = 1e-4 if not noise else 1e-3      # with noise on, ulp differences of logf/cosf in Box-Muller enter the loop as 1e-7-level input noise
    steps = g["state_steps"]
    sel = steps <= T_cmp
    close_relative(out["states"][steps[sel]], g["states"][sel], rtol, STATE_GROUPS, "states")
    close_relative(out["actions"][:T_cmp], g["actions"][:T_cmp], rtol, {"action": (slice(0, 4), 0.1)}, "actions")
    close(out["rewards"][:T_cmp], g["rewards"][:T_cmp], 1e-3, 1e-3, "rewards")
    assert np.array_equal(out["terminated"], g["terminated"])
    assert np.array_equal(env.get_rng(), g["final_rng"])
    h, gs = env.get_hidden()
    assert np.array_equal(gs, g["final_gru_step"])
    # the whole horizon stays bounded-close as well (closed loop is stable under the policy)
    close_relative(out["states"][steps], g["states"], 50 * rtol, STATE_GROUPS, "states, full horizon")
    close(env.get_state(), out["states"][-1], 0, 0, "final state row == slot 0")


@pytest.mark.parametrize("spec", [B.SPEC_DEFAULT, B.SPEC_RAPTOR])
def test_fused_rollout_equals_stepwise_api(rb, port, spec):
    """ONE fused launch == T x (observe, evaluate_step, step) through the vector API"""
    n, T = 300, 40
    blob = rb.raptor_policy_blob()
    a_env = rb.VectorEnvironment(n, spec)
    b_env = rb.VectorEnvironment(n, spec)
    for e in (a_env, b_env):
        e.initialize_rng(11, warmup=20)
        e.sample_initial_state()
        e.load_policy(blob)
    out = a_env.rollout(T, record=("actions", "rewards", "terminated", "returns", "episode_length"))
    acts, rews = [], []
    for t in range(T):
        o = b_env.observe()
        a = b_env.policy_evaluate_step(np.ascontiguousarray(o[:, :22]))
        b_env.step(a)
        rews.append(b_env.reward(a))
        b_env.copy_state(0, 1)
        acts.append(a.copy())
    # same device functions, different kernels: the compiler may contract different multiply-adds into FMAs, so compare at 1e-4 relative
    close_relative(out["actions"], np.array(acts), 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")
    close(out["rewards"], np.array(rews), 1e-3, 1e-3, "rewards")
    close_relative(a_env.get_state()[None], b_env.get_state()[None], 1e-4, STATE_GROUPS, "final states")
    assert np.array_equal(a_env.get_rng(), b_env.get_rng())
    # evaluate()-style statistics: return accumulated until (and including) the first terminated step
    term = out["terminated"].astype(bool)
    first = np.where(term.any(0), term.argmax(0), T - 1)
    want_len = first + 1
    assert np.array_equal(out["episode_length"], want_len)
    want_ret = np.array([out["rewards"][:want_len[i], i].sum() for i in range(n)])
    close(out["returns"], want_ret, 1e-4, 1e-3, "returns")


def test_full_size_properties(rb):
    """BASELINE config 2 size (65 536 envs): determinism, shard invariance (global env ids), quaternion norm, bounded states"""
    n, T = 65536, 200
    blob = rb.raptor_policy_blob()

    def run(n_envs, first):
        e = rb.VectorEnvironment(n_envs, rb.SPEC_RAPTOR, first_env_id=first)
        e.initialize_rng(2024, warmup=16)
        e.sample_initial_state()
        e.load_policy(blob)
        o = e.rollout(T, record=("returns", "episode_length"))
        return e.get_state(), o["returns"], o["episode_length"], e.get_rng()

    s1, r1, l1, g1 = run(n, 0)
    s2, r2, l2, g2 = run(n, 0)
    assert np.array_equal(s1, s2) and np.array_equal(r1, r2) and np.array_equal(g1, g2)       # deterministic
    sa, ra, la, ga = run(n // 2, 0)
    sb, rb_, lb, gb = run(n // 2, n // 2)
    assert np.array_equal(np.concatenate([sa, sb]), s1) and np.array_equal(np.concatenate([ra, rb_]), r1)   # shard invariant
    assert np.array_equal(np.concatenate([ga, gb]), g1)
    q = s1[:, 3:7]
    close(np.linalg.norm(q, axis=1), np.ones(n), 1e-5, 0, "unit quaternions")
    assert np.isfinite(s1).all()
    alive = l1 == T
    assert alive.mean() > 0.8            # the policy keeps the nominal Crazyflie in the box from the init distribution
    assert np.abs(s1[alive, :3]).max() <= 1.0 + 1e-6


@pytest.mark.parametrize("gemm", ["tcgen05", "fp32"])
def test_no_auto_reset_mode(rb, port, gemm):
    """nn::layers::gru::NoAutoResetMode (gru/operations_generic.h:343-411 with the mode set): the step counter keeps counting past SEQUENCE_LENGTH
    and the hidden state is never re-initialised; fused rollout across the 500-step boundary against the oracle, then the evaluate_step path"""
    n, T = 200, 520
    spec = rb.SPEC_RAPTOR
    env = rb.VectorEnvironment(n, spec)
    env.initialize_rng(77, warmup=16)
    env.sample_initial_state()
    env.load_policy(gemm=rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32)
    params, states, rng = env.get_parameters(), env.get_state(), env.get_rng()
    out = env.rollout(T, record=("actions", "terminated"), no_auto_reset=True)
    pol = port.make_policy(rb.raptor_policy_blob())
    hid, gs = np.tile(np.load(os.path.join(G, "raptor_kat.npz"))["h0"], (n, 1)).astype(np.float32), np.zeros(n, np.int32)
    want = port.rollout(spec, pol, params, states, rng, T, hidden=hid, gru_step=gs, no_auto_reset=True)
    h, g = env.get_hidden()
    assert np.all(g == T) and np.array_equal(g, gs)                                 # no wrap at 500
    assert np.array_equal(env.get_rng(), rng)
    close_relative(out["actions"][:100], want["actions"][:100], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")
    close(out["actions"][495:], want["actions"][495:], 5e-3, 5e-3, "actions across the SEQUENCE_LENGTH boundary")
    close(h, hid, 5e-3, 5e-3, "hidden state")
    assert not np.allclose(h, np.tile(np.load(os.path.join(G, "raptor_kat.npz"))["h0"], (n, 1)), atol=1e-3)   # it was NOT reset
    # the same rollout WITH auto-reset differs after the boundary (the mode really switches something)
    env2 = rb.VectorEnvironment(n, spec)
    env2.set_parameters(params); env2.initialize_rng(77, warmup=16); env2.sample_initial_state(); env2.load_policy()
    out2 = env2.rollout(T, record=("actions",))
    assert np.abs(out2["actions"][505:] - out["actions"][505:]).max() > 1e-3
    assert np.all(env2.get_hidden()[1] == T - 500)
    # evaluate_step path: 3 steps with the flag, counter keeps counting from an arbitrary value beyond the sequence length
    env.set_hidden(gru_step=np.full(n, 499, np.int32))
    obs = np.ascontiguousarray(env.observe()[:, :22])
    hid2, gs2 = env.get_hidden()[0].copy(), np.full(n, 499, np.int32)
    for _ in range(3):
        a = env.policy_evaluate_step(obs, no_auto_reset=True)
        a_want, _, _ = port.policy_evaluate_step(pol, obs, hidden=hid2, gru_step=gs2, no_auto_reset=True)
        close(a, a_want, 1e-4, 1e-5, "evaluate_step, NoAutoResetMode")
    assert np.all(env.get_hidden()[1] == 502) and np.all(gs2 == 502)


def test_bench_configuration_vs_oracle(rb, port):
    """the exact launch bench.py times (BASELINE configs[1]): SPEC_RAPTOR_DR, 65 536 environments, foundation-policy DR ranges, default math, tcgen05,
    default kernel selection and time-chunk schedule; actions / states of a strided 512-environment sample over 100 steps against the oracle at the
    north-star bound (1e-4 relative), terminated flags, RNG streams and GRU counters bit-exact"""
    import bench
    n, T = 65536, 100
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
    row = env.get_environment_parameters()
    row[124:139] = np.array(bench.DR_RANGES, np.float32)
    env.set_environment_parameters(row)
    env.initialize_rng(seed=20250925, warmup=16)
    env.sample_initial_parameters()
    env.sample_initial_state()
    env.load_policy(gemm=rb.GEMM_TCGEN05_3XTF32)
    sel = np.arange(0, n, 128) + (np.arange(n // 128) * 37) % 128          # one environment of every 128-row tile, varying lane
    params, states, rng = env.get_parameters()[sel], env.get_state()[sel], env.get_rng()[sel]
    out = env.rollout(T, record=("actions", "terminated", "rewards", "states"), state_stride=10)
    pol = port.make_policy(rb.raptor_policy_blob())
    hid, gs = np.tile(np.load(os.path.join(G, "raptor_kat.npz"))["h0"], (len(sel), 1)).astype(np.float32), np.zeros(len(sel), np.int32)
    want = port.rollout(rb.SPEC_RAPTOR_DR, pol, params, states, rng, T, hidden=hid, gru_step=gs)
    close_relative(out["actions"][:, sel], want["actions"], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")
    close(out["rewards"][:, sel], want["rewards"], 1e-3, 1e-3, "rewards")
    assert np.array_equal(out["terminated"][:, sel], want["terminated"])
    assert np.array_equal(env.get_rng()[sel], rng)
    h, g = env.get_hidden()
    assert np.array_equal(g[sel], gs)
    close(h[sel], hid, 1e-3, 1e-4, "hidden state")
    # states every 10 steps, |error| relative to the magnitude of each physical quantity along the trajectory (the rule of test_fused_rollout_vs_golden)
    close_relative(out["states"][:, sel], want["states"][::10], 1e-4, STATE_GROUPS, "states")
    close(env.get_state()[sel], out["states"][-1][:, :][sel], 0, 0, "final state row == slot 0")


@pytest.mark.gpu
def test_config3_and_config4_launches_vs_oracle(rb, port):
    """the other two single-GPU launches bench.py times, at FULL size, against the oracle on a strided environment sample (one environment of every eighth tile,
    varying lane): config 3 = 1 048 576 environments, per-environment DR, SAC-teacher MLP 26-64-64-8 (k_rollout_mlp_ts, time-chunked scheduler, 2 CTAs/SM);
    config 4 = 262 144 environments, PPO collection with in-kernel resets and the bulk-copy write-back (k_collect_ts).  First 20 steps at the north-star bound,
    flags and RNG streams bit-exact."""
    import bench
    import torch
    rs = np.random.RandomState(0)
    dev = torch.device("cuda", 0)
    # ---- config 3
    n, T = 1048576, 20
    env = rb.VectorEnvironment(n, rb.SPEC_TEACHER_DR)
    row = env.get_environment_parameters(); row[124:139] = np.array(bench.DR_RANGES, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(3, warmup=16); env.sample_initial_parameters(); env.sample_initial_state()
    blob = bench.mlp_blob(rs, 26, 8, False, False)
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=rb.GEMM_TCGEN05_3XTF32)
    sel = np.arange(0, n, 1024) + (np.arange(n // 1024) * 37) % 128
    params, states, rng = env.get_parameters()[sel], env.get_state()[sel], env.get_rng()[sel]
    out = dict(actions=torch.zeros((T, n, 4), device=dev), rewards=torch.zeros((T, n), device=dev), terminated=torch.zeros((T, n), dtype=torch.uint8, device=dev))
    env.rollout(T, out=out)
    env.synchronize()
    assert env.last_kernel() == "k_rollout_mlp_ts"
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_EVAL)
    want = port.rollout(rb.SPEC_TEACHER_DR, pol, params, states, rng, T)
    tsel = torch.from_numpy(sel).to(dev)
    close_relative(out["actions"][:, tsel].cpu().numpy(), want["actions"], 1e-4, {"action": (slice(0, 4), 0.1)}, "config 3 actions")
    close(out["rewards"][:, tsel].cpu().numpy(), want["rewards"], 1e-3, 1e-3, "config 3 rewards")
    assert np.array_equal(out["terminated"][:, tsel].cpu().numpy(), want["terminated"])
    assert np.array_equal(env.get_rng()[sel], rng)
    close_relative(env.get_state()[sel][None], states[None], 1e-4, STATE_GROUPS, "config 3 final states")
    del env, out
    # ---- config 4
    n, T, limit = 262144, 20, 500
    obs, D = 22, 37
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
    row = env.get_environment_parameters(); row[124:139] = np.array(bench.DR_RANGES, np.float32); env.set_environment_parameters(row)
    env.initialize_rng(4, warmup=16); env.initial_parameters(); env.initial_state()
    blob = bench.mlp_blob(rs, 22, 4, True, True)
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=rb.GEMM_TCGEN05_3XTF32)
    env.collect_reset()
    sel = np.arange(0, n, 1024) + (np.arange(n // 1024) * 37) % 128
    params, states, rng = env.get_parameters()[sel], env.get_state()[sel], env.get_rng()[sel]
    data = torch.full(((T + 1) * n, D), 7.0, dtype=torch.float32, device=dev)          # stale contents: every row must be rewritten
    env.collect(T, limit, data)
    env.synchronize()
    assert env.last_kernel() == "k_collect_ts"
    got3 = data.view(T + 1, n, D)[:, torch.from_numpy(sel).to(dev)].cpu().numpy()
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=B.HEAD_PPO_GAUSSIAN)
    m = len(sel)
    ep_step = np.zeros(m, np.int32); ep_ret = np.zeros(m, np.float32); trunc = np.ones(m, np.uint8)
    want3 = port.collect(rb.SPEC_RAPTOR_DR, pol, row, params, states, rng, ep_step, ep_ret, trunc, T, limit).reshape(T + 1, m, D)
    assert np.array_equal(env.get_rng()[sel], rng)
    assert np.array_equal(got3[:T, :, obs + 10:obs + 12], want3[:T, :, obs + 10:obs + 12])      # terminated | truncated
    assert want3[:T, :, obs + 11].sum() > 0                                                      # episodes ended (and were reset in the kernel) inside the window
    K = 8
    close(got3[:K, :, :obs], want3[:K, :, :obs], 1e-4, 3e-5, "config 4 observations, first steps")
    close(got3[:K, :, obs:obs + 8], want3[:K, :, obs:obs + 8], 1e-4, 3e-5, "config 4 action means / actions, first steps")
    close(got3[..., :obs], want3[..., :obs], 2e-3, 2e-4, "config 4 observations (incl. the final row)")
    close(got3[:T, :, obs + 8], want3[:T, :, obs + 8], 1e-3, 1e-3, "config 4 log-prob")
    assert np.all(got3[..., obs + 12:] == 0) and np.all(got3[T, :, obs:] == 0)                   # learner columns and the final rows' step columns: zeros
    assert int((data == 7.0).sum().item()) == 0                                                  # no stale element survived anywhere in the 0.8 GB dataset
    close(env.get_parameters()[sel], params, 2e-6, 0, "config 4 parameters after the in-kernel resets (deferred write-back)")


def test_asynchronous_transfers_match_the_synchronous_calls(rb):
    """b200l2f_set_parameters_async / set_state_async / get_state_async / copy_to_host_async (copy streams + staging buffers, transposes ordered on the
    main stream): a pipelined sequence of rollouts fed from page-locked host memory gives the bits of the synchronous calls; pageable memory falls back"""
    import torch
    n, T = 3000, 50
    def make():
        e = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
        e.initialize_rng(5, warmup=16)
        return e
    env = make()
    row = env.get_environment_parameters()
    import bench
    row[124:139] = np.array(bench.DR_RANGES, np.float32)
    env.set_environment_parameters(row)
    env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy()
    p0, s0, r0 = env.get_parameters(), env.get_state(), env.get_rng()
    ref = make()
    ref.load_policy()
    want_states, want_returns = [], []
    for k in range(3):                      # step k starts from s0 shifted by k rows: every step has its own inputs
        ref.set_parameters(np.roll(p0, k, axis=0)); ref.set_state(np.roll(s0, k, axis=0)); ref.set_rng(r0); ref.policy_reset()
        o = ref.rollout(T, record=("returns",))
        want_states.append(ref.get_state()); want_returns.append(o["returns"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    tp = [pin(np.roll(p0, k, axis=0)) for k in range(3)]; ts = [pin(np.roll(s0, k, axis=0)) for k in range(3)]
    hs = [pin(np.zeros_like(s0)) for _ in range(3)]; hr = [pin(np.zeros(n, np.float32)) for _ in range(3)]
    rd = [torch.zeros(n, dtype=torch.float32, device="cuda") for _ in range(3)]
    env.set_parameters_async(tp[0].numpy()); env.set_state_async(ts[0].numpy())
    for k in range(3):
        env.set_rng(r0); env.policy_reset()
        env.rollout(T, out={"returns": rd[k]})
        env.get_state_async(hs[k].numpy()); env.copy_to_host_async(hr[k].numpy(), rd[k])
        if k + 1 < 3:
            env.set_parameters_async(tp[k + 1].numpy()); env.set_state_async(ts[k + 1].numpy())
    env.transfers_synchronize(); env.synchronize()
    for k in range(3):
        assert np.array_equal(hs[k].numpy(), want_states[k]), k
        assert np.array_equal(hr[k].numpy(), want_returns[k]), k
    # pageable host memory: the same calls work, synchronously
    env.set_parameters_async(p0); env.set_state_async(s0)
    got = env.get_state_async(np.zeros_like(s0))
    assert np.array_equal(got, s0) and np.array_equal(env.get_parameters(), p0)
    with pytest.raises(ValueError):
        env.set_state_async(s0[:, :10])


def test_last_status_aggregates_and_nonfinite_flags(rb):
    """b200l2f_last_status: the rl::utils::evaluation::Result aggregates (operations_generic.h:201-213) of a rollout reduced on the device, and the
    per-environment non-finite flag (05_state_is_nan.h) with its count"""
    n, T = 1000, 120
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR)
    env.initialize_rng(3, warmup=16)
    p = env.get_parameters(); p[:, 115] = 0.6; env.set_parameters(p)        # tighter position threshold: some episodes terminate
    env.sample_initial_state(); env.load_policy()
    out = env.rollout(T, record=("returns", "episode_length", "terminated"))
    st = env.last_status(flags=True)
    term = out["terminated"].any(0)
    assert st["n_envs"] == n and st["has_episodes"] == 1 and st["n_nonfinite"] == 0 and not st["nonfinite_flags"].any()
    assert st["n_terminated"] == int(term.sum()) and 0 < term.sum() < n
    np.testing.assert_allclose(st["share_terminated"], term.mean(), rtol=1e-12)
    np.testing.assert_allclose(st["returns_mean"], out["returns"].astype(np.float64).mean(), rtol=1e-9)
    np.testing.assert_allclose(st["returns_std"], out["returns"].astype(np.float64).std(), rtol=1e-7)
    np.testing.assert_allclose(st["episode_length_mean"], out["episode_length"].mean(), rtol=1e-12)
    np.testing.assert_allclose(st["episode_length_std"], out["episode_length"].astype(np.float64).std(), rtol=1e-9)
    # a rollout that does not ask for returns still reports them
    env.sample_initial_state(); env.policy_reset()
    env.rollout(T)
    assert env.last_status()["has_episodes"] == 1 and np.isfinite(env.last_status()["returns_mean"])
    # poisoned states are flagged and counted
    s = env.get_state(); s[7, 0] = np.nan; s[400, 9] = np.nan; env.set_state(s)       # (an Inf would be clamped to the +-1e5 state limit by the step)
    env.rollout(3)
    st = env.last_status(flags=True)
    assert st["n_nonfinite"] >= 2 and st["nonfinite_flags"][7] == 1 and st["nonfinite_flags"][400] == 1 and st["nonfinite_flags"].sum() == st["n_nonfinite"]
    fresh = rb.VectorEnvironment(4, rb.SPEC_RAPTOR)
    with pytest.raises(rb.EngineError, match="no rollout"):
        fresh.last_status()


@pytest.mark.parametrize("spec", [B.SPEC_DEFAULT, B.SPEC_RAPTOR, B.SPEC_TEACHER])
def test_step_repeated_equals_repeated_steps(rb, port, spec):
    """b200l2f_step_repeated (the loop of the reference's GPU benchmark, benchmark.cu:111-120: T x step under a held action, state in registers) against
    T x rl_tools::step of the oracle from the same states: 1e-4 relative over 50 steps, RNG streams (Langevin targets) bit-exact"""
    n, T = 300, 50
    env = rb.VectorEnvironment(n, spec)
    env.initialize_rng(9, warmup=16)
    env.sample_initial_state()
    p, s0, rng = env.get_parameters(), env.get_state(), env.get_rng()
    a = np.array([0.1, -0.2, 0.05, 0.3], np.float32)
    env.step_repeated(a, T)
    got = env.get_state()
    want = s0.copy()
    for i in range(n):
        r = rng[i:i + 1].copy()
        for _ in range(T):
            want[i], _dt = port.step(spec, p[i], want[i], a, r)
        rng[i] = r[0]
    assert np.array_equal(env.get_rng(), rng)
    close_relative(got[None], want[None], 1e-4, STATE_GROUPS, "states after %d held-action steps" % T)
    env.step_repeated(a, 0)
    assert np.array_equal(env.get_state(), got)


def test_ragged_sizes_and_errors(rb):
    for n in [1, 31, 129, 1000]:
        e = rb.VectorEnvironment(n, rb.SPEC_DEFAULT)
        e.initialize_rng(1, warmup=8)
        e.sample_initial_state()
        e.load_policy()
        o = e.rollout(3, record=("actions",))
        assert o["actions"].shape == (3, n, 4) and np.isfinite(o["actions"]).all()
        o0 = e.rollout(0, record=("returns",))
        assert np.all(o0["returns"] == 0)
    e = rb.VectorEnvironment(4, rb.SPEC_RAPTOR)
    with pytest.raises(rb.EngineError):
        e.rollout(1)                                 # no policy loaded
    with pytest.raises(rb.EngineError):
        e.observe(slot=7)                            # bad slot
    with pytest.raises(rb.EngineError):
        bad = e.get_environment_parameters(); bad[126] = 1.0; e.set_environment_parameters(bad); e.sample_initial_parameters()  # DR range without DR spec
    with pytest.raises(rb.EngineError):
        rb.VectorEnvironment(0)


@pytest.mark.parametrize("gemm", ["tcgen05", "fp32"])
def test_runner_edge_cases(rb, port, gemm):
    """empty and minimal inputs of the runner-style entry points: zero steps leave everything untouched, one environment / ragged tiles / a ring of
    one row work, and the collection with zero steps writes just the final observations"""
    g = rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32
    rs = np.random.RandomState(3)
    for n in (1, 33):
        # ---- PPO collection
        env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR)
        env.initialize_rng(5, warmup=8); env.initial_parameters(); env.initial_state()
        env.load_policy(random_mlp_blob(rs, 22, 4, True, True), arch=rb.POLICY_MLP, input_dim=22, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=g)
        env.collect_reset()
        rng0, s0 = env.get_rng(), env.get_state()
        d0 = env.collect(0, 10)
        assert d0.shape == (n, 37) and np.all(d0[:, 22:] == 0) and np.isfinite(d0).all()
        assert np.array_equal(env.get_rng(), rng0) and np.array_equal(env.get_state(), s0)       # T = 0: only the final observe (noise-free), no reset
        d1 = env.collect(1, 1)                                                                    # step limit 1: every step truncates
        assert d1.shape == (2 * n, 37) and np.all(d1[:n, 33] == 1)
        # ---- off-policy runner
        env = rb.VectorEnvironment(n, rb.SPEC_TEACHER)
        env.initialize_rng(6, warmup=8); env.initial_parameters(); env.initial_state()
        env_blob = random_mlp_blob(rs, 26, 8, False, False)
        env.load_policy(env_blob, arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=g)
        env.collect_reset()
        ring = env.new_replay_buffers(1)                                                          # capacity 1: every add wraps
        rng0 = env.get_rng()
        env.off_policy_steps(0, 5, ring)
        assert np.array_equal(env.get_rng(), rng0) and not ring["full"].any() and np.all(ring["data"] == 0)
        env.off_policy_steps(3, 5, ring)
        assert ring["full"].all() and np.all(ring["position"] == 0) and np.all(ring["episode_start"] == 0) and np.isfinite(ring["data"]).all()
        step, ret, trunc = env.get_runner_state()
        assert np.all(step == 3) and np.all(trunc == 0)
        b = env.gather_batch(ring, port.rng_states(1, 5, warmup=3), 500)
        assert np.all(b["sample_index"] == 0) and np.array_equal(b["observations_actions"][0, :, :26], ring["data"][b["env_index"], 0, :26])
        # regression: fully inactive warps of the last tile (their lanes shadow environment 0) must not write ring rows -- once an unsigned
        # `n - warp_env0` made them stream their un-reset copy of environment 0 over its rows (timing-dependent; found with compute-sanitizer)
        for rep in range(4):
            e2 = rb.VectorEnvironment(n, rb.SPEC_TEACHER)
            e2.initialize_rng(60 + rep, warmup=8); e2.initial_parameters(); e2.initial_state()
            e2.load_policy(env_blob, arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=g)
            e2.collect_reset()
            r2 = e2.new_replay_buffers(8)
            rng2, p2, s2 = e2.get_rng(), e2.get_parameters(), e2.get_state()
            e2.off_policy_steps(6, 50, r2, sample_parameters=False)
            run2 = B.new_off_policy_runner(n, 8, 26)
            port.off_policy_steps(B.SPEC_TEACHER, port.make_policy(env_blob, arch=B.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_SAMPLE),
                                  port.nominal_parameters(B.SPEC_TEACHER), p2, s2, rng2, run2, 6, 50, sample_parameters=False)
            close(r2["data"][..., :26], run2["replay"][..., :26], 2e-3, 2e-4, "ring observations, n = %d" % n)
        env.set_runner_state(truncated=np.ones(n, np.uint8))                                      # truncate_all (operations_generic.h:152-154)
        env.off_policy_steps(1, 5, ring)
        assert np.all(env.get_runner_state()[0] == 1)
        with pytest.raises(ValueError):
            env.off_policy_steps(1, 5, env.new_replay_buffers(1) | {"data": np.zeros((n, 1, 58), np.float32)})


@pytest.mark.parametrize("name,T_cmp", [("default_8x500.npz", 100), ("raptor_dr_64x100.npz", 100)])
def test_tcgen05_rollout_vs_golden(rb, name, T_cmp):
    """actor GEMMs on the tensor cores (tcgen05, 3xTF32): same 1e-4 closed-loop bound against the reference fixtures"""
    g = np.load(os.path.join(G, name))
    spec = int(g["spec"])
    n, T = g["states0"].shape[0], g["actions"].shape[0]
    env = rb.VectorEnvironment(n, spec)
    env.set_parameters(np.ascontiguousarray(g["params"]))
    env.set_state(np.ascontiguousarray(g["states0"]))
    env.set_rng(np.ascontiguousarray(g["rng0"]))
    env.load_policy(np.load(os.path.join(G, "raptor_kat.npz"))["blob"], gemm=rb.GEMM_TCGEN05_3XTF32)
    out = env.rollout(T, record=("states", "actions", "rewards", "terminated"))
    steps = g["state_steps"]
    sel = steps <= T_cmp
    close_relative(out["states"][steps[sel]], g["states"][sel], 1e-4, STATE_GROUPS, "states")
    close_relative(out["actions"][:T_cmp], g["actions"][:T_cmp], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")
    assert np.array_equal(out["terminated"], g["terminated"])
    assert np.array_equal(env.get_rng(), g["final_rng"])
    assert np.array_equal(env.get_hidden()[1], g["final_gru_step"])


def test_tcgen05_rollout_matches_fp32_rollout(rb):
    """same inputs through both GEMM back ends, 3000 envs (ragged last CTA), 600 steps (crosses the GRU auto-reset at 500)"""
    n, T = 3000, 600
    outs = []
    for gemm in (rb.GEMM_FP32_CUDA_CORES, rb.GEMM_TCGEN05_3XTF32):
        e = rb.VectorEnvironment(n, rb.SPEC_RAPTOR)
        e.initialize_rng(77, warmup=16)
        e.sample_initial_state()
        e.load_policy(gemm=gemm)
        o = e.rollout(T, record=("actions", "returns", "episode_length"), )
        outs.append((o, e.get_state(), e.get_hidden(), e.get_rng()))
    (oa, sa, (ha, ga), ra), (ob, sb, (hb, gb), rb_) = outs
    assert np.array_equal(ra, rb_) and np.array_equal(ga, gb)
    stable = (oa["episode_length"] == T) & (ob["episode_length"] == T)
    assert stable.mean() > 0.8
    close_relative(ob["actions"][:100, stable], oa["actions"][:100, stable], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions, first 100 steps")
    close_relative(sb[None, stable], sa[None, stable], 5e-3, STATE_GROUPS, "final states after 600 steps")
    assert (oa["episode_length"] == ob["episode_length"]).mean() > 0.995


@pytest.mark.parametrize("actor", ["raptor_gru", "teacher_mlp"])
def test_tcgen05_noise_variants_match_cuda_core_kernels(rb, port, actor):
    """observation / action noise inside the tcgen05 rollout kernels (MUFU Box-Muller) against the CUDA-core kernels (libdevice Box-Muller) and
    the oracle: the integer RNG streams are bit-exact, trajectories agree to the closed-loop tolerance of the noise fixture"""
    n, T = 300, 60
    spec = rb.SPEC_RAPTOR if actor == "raptor_gru" else rb.SPEC_TEACHER
    row = port.nominal_parameters(spec).copy()
    row[108:114] = np.array([0.02, 0.03, 0.05, 0.1, 0.0, 0.05], np.float32)   # imu noise unused by these observations; action noise on
    blob = None if actor == "raptor_gru" else random_mlp_blob(np.random.RandomState(9), 26, 8, False, False)
    outs = []
    for gemm in (rb.GEMM_FP32_CUDA_CORES, rb.GEMM_TCGEN05_3XTF32):
        e = rb.VectorEnvironment(n, spec)
        e.set_environment_parameters(row)
        e.initialize_rng(21, warmup=16)
        e.initial_parameters()
        e.sample_initial_state()
        if actor == "raptor_gru":
            e.load_policy(gemm=gemm)
        else:
            e.load_policy(blob, arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=gemm)
        params, states, rng = e.get_parameters(), e.get_state(), e.get_rng()
        o = e.rollout(T, record=("observations", "actions", "terminated"))
        outs.append((o, e.get_rng()))
    (oa, ra), (ob, rb_) = outs
    assert np.array_equal(ra, rb_)
    close(ob["observations"][:20], oa["observations"][:20], 1e-3, 1e-4, "noisy observations, first 20 steps")
    close(ob["actions"][:20], oa["actions"][:20], 2e-3, 2e-3, "actions, first 20 steps")
    assert (oa["terminated"] == ob["terminated"]).mean() > 0.999
    # and against the oracle (same stream, same draw order)
    if actor == "raptor_gru":
        pol = port.make_policy(rb.raptor_policy_blob())
        h = np.tile(rb.raptor_policy_blob()[352 + 16 + 2 * (768 + 48):][:16], (n, 1)).astype(np.float32)
        want = port.rollout(spec, pol, params, states, rng, T, hidden=h, gru_step=np.zeros(n, np.int32))
    else:
        pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_EVAL)
        want = port.rollout(spec, pol, params, states, rng, T)
    assert np.array_equal(rng, rb_)
    obs_dim = 22 if actor == "raptor_gru" else 26
    close(ob["observations"][:20], want["observations"][:20, :, :obs_dim], 1e-3, 1e-4, "noisy observations vs oracle")
    close(ob["actions"][:20], want["actions"][:20], 2e-3, 2e-3, "actions vs oracle")


def test_per_environment_mdp_parameters(rb, port):
    """reward weights / termination thresholds that differ per environment take the general (non-uniform) kernel path"""
    n, T = 256, 30
    spec = rb.SPEC_RAPTOR
    rs = np.random.RandomState(3)
    params = np.tile(port.nominal_parameters(spec), (n, 1))
    params[:, 96] = rs.uniform(0.2, 2.0, n)        # reward.constant
    params[:, 98] = rs.uniform(0.5, 1.5, n)        # reward.position
    params[:, 101] = rs.uniform(0.0, 0.3, n)       # reward.linear_velocity
    params[:, 116] = rs.uniform(1.0, 3.0, n)       # termination.linear_velocity_threshold
    params = params.astype(np.float32)
    rng = port.rng_states(9, n, warmup=16)
    states = port.sample_initial_state_n(spec, params, rng)
    pol = port.make_policy(rb.raptor_policy_blob())
    h = np.tile(rb.raptor_policy_blob()[352 + 16 + 2 * (768 + 48):][:16], (n, 1)).astype(np.float32)
    want = port.rollout(spec, pol, params, states.copy(), rng.copy(), T, hidden=h, gru_step=np.zeros(n, np.int32))
    for gemm in (rb.GEMM_FP32_CUDA_CORES, rb.GEMM_TCGEN05_3XTF32):
        e = rb.VectorEnvironment(n, spec)
        e.set_parameters(params); e.set_state(states); e.set_rng(rng)
        e.load_policy(gemm=gemm)
        out = e.rollout(T, record=("rewards", "terminated", "actions"))
        assert np.array_equal(out["terminated"], want["terminated"])
        close(out["rewards"], want["rewards"], 1e-4, 2e-4, "rewards")
        close_relative(out["actions"], want["actions"], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")


# ---------------------------------------------------------------------------------------------------------------------------
# MLP actors (SAC teacher, PPO) and PPO collection
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec,head", [(B.SPEC_RAPTOR, "ppo"), (B.SPEC_TEACHER, "squash"), (B.SPEC_DEFAULT, "identity")])
def test_mlp_evaluate_step(rb, port, spec, head):
    n = 300
    rs = np.random.RandomState(7)
    in_dim = port.observation_dim(spec)
    out_dim = 8 if head == "squash" else 4
    hd_id = {"identity": (rb.HEAD_IDENTITY, B.HEAD_IDENTITY), "squash": (rb.HEAD_SQUASH_EVAL, B.HEAD_SQUASH_EVAL), "ppo": (rb.HEAD_PPO_GAUSSIAN, B.HEAD_PPO_GAUSSIAN)}[head]
    blob = random_mlp_blob(rs, in_dim, out_dim, True, head == "ppo")
    env = rb.VectorEnvironment(n, spec)
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=in_dim, hidden_dim=64, output_dim=out_dim, standardize=1, head=hd_id[0])
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=in_dim, hidden_dim=64, output_dim=out_dim, standardize=1, head=hd_id[1])
    obs = rs.normal(0, 1, (n, in_dim)).astype(np.float32)
    rng = port.rng_states(4, n, warmup=16)
    env.set_rng(rng)
    got = env.policy_evaluate_step(obs)
    want, _, _ = port.policy_evaluate_step(pol, obs, rng=rng)
    close(got, want, 2e-5, 2e-5, head)
    if head == "ppo":
        assert np.array_equal(env.get_rng(), rng)


@pytest.mark.parametrize("gemm", ["fp32", "tcgen05"])
def test_rollout_with_teacher_mlp(rb, port, gemm):
    """BASELINE config 3 shape: TEACHER spec (OBS 26), per-env DR dynamics, SAC-teacher MLP 26-64-64-8 + squash (evaluation mode);
    CUDA-core kernel (k_rollout_mlp) and tensor-core kernel (k_rollout_mlp_ts) against the same oracle rollout"""
    gemm = rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32
    n, T = 200, 60
    spec = B.SPEC_TEACHER_DR
    rs = np.random.RandomState(11)
    blob = random_mlp_blob(rs, 26, 8, False, False)
    env_p = foundation_dr_env_params(port, spec)
    rng = port.rng_states(21, n, warmup=16)
    params = port.sample_initial_parameters_n(spec, env_p, rng)
    states = port.sample_initial_state_n(spec, params, rng)
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_EVAL)
    want = port.rollout(spec, pol, params, states.copy(), rng.copy(), T)
    env = rb.VectorEnvironment(n, spec)
    env.set_parameters(params); env.set_state(states); env.set_rng(rng)
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=26, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=gemm)
    out = env.rollout(T, record=("states", "observations", "actions", "rewards", "terminated"))
    # a random actor does not stabilise the vehicle: compare while trajectories are still regular (first 20 steps) at 1e-4, the rest loosely
    close_relative(out["actions"][:20], want["actions"][:20], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions")
    close_relative(out["states"][:21], want["states"][:21], 1e-4, STATE_GROUPS, "states")
    close(out["observations"][:20], want["observations"][:20], 1e-3, 1e-4, "observations")
    assert (out["terminated"][:20] == want["terminated"][:20]).mean() > 0.999
    assert np.array_equal(env.get_rng(), port_final_rng(port, spec, pol, params, states, rng, T))


def port_final_rng(port, spec, pol, params, states, rng, T):
    r = rng.copy()
    port.rollout(spec, pol, params, states.copy(), r, T, record=False)
    return r


@pytest.mark.parametrize("spec,gemm", [(s_, g_) for s_ in (B.SPEC_RAPTOR, B.SPEC_RAPTOR_DR) for g_ in ("fp32", "tcgen05", "tcgen05-edited-parameters")]
                         + [(B.SPEC_DEFAULT, "fp32"), (B.SPEC_DEFAULT_DR, "fp32"), (B.SPEC_DEFAULT_DR, "fp32-edited-parameters"),
                            (B.SPEC_DEFAULT, "tcgen05"), (B.SPEC_DEFAULT_DR, "tcgen05"), (B.SPEC_DEFAULT_DR, "tcgen05-edited-parameters")])
def test_ppo_collect_vs_oracle(rb, port, spec, gemm):
    """BASELINE config 4 shape: PPO actor (standardize -> 64 -> 64 -> 4, learned log_std), Gaussian sampling, auto-reset on
    terminated-or-step-limit with re-sampled parameters and state, dataset rows in the reference layout; CUDA-core kernel (k_collect)
    and tensor-core kernel (k_collect_ts); 200 environments = one full tile + a ragged one.  DEFAULT specs: the PPO zoo's environment
    (rl/zoo/l2f/ppo.h: H = 16 action history, 82-wide observation, 97-float rows) on the CUDA-core kernel."""
    edited = gemm.endswith("edited-parameters")   # parameters written by the caller: the kernel may not assume the columns follow the nominal row
    gemm = rb.GEMM_FP32_CUDA_CORES if gemm.startswith("fp32") else rb.GEMM_TCGEN05_3XTF32
    n, T, limit = 200, 40, 12
    obs = port.observation_dim(spec)
    rs = np.random.RandomState(5)
    blob = random_mlp_blob(rs, obs, 4, True, True)
    env = rb.VectorEnvironment(n, spec)
    env_p = foundation_dr_env_params(port, spec) if spec in (B.SPEC_RAPTOR_DR, B.SPEC_DEFAULT_DR) else port.nominal_parameters(spec)
    env.set_environment_parameters(env_p)
    env.initialize_rng(31, warmup=16)
    env.initial_parameters()
    env.initial_state()
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=gemm)
    env.collect_reset()
    rng = env.get_rng()
    params, states = env.get_parameters(), env.get_state()
    if edited:   # per-environment values OUTSIDE the domain-randomised set (reward scale / constant): in effect until the environment's first reset
        params[:, 95] *= rs.uniform(0.5, 1.5, n).astype(np.float32)
        params[:, 96] += rs.uniform(0.0, 0.2, n).astype(np.float32)
        env.set_parameters(params)
    data = env.collect(T, limit)
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=B.HEAD_PPO_GAUSSIAN)
    ep_step = np.zeros(n, np.int32); ep_ret = np.zeros(n, np.float32); trunc = np.ones(n, np.uint8)
    want = port.collect(spec, pol, env_p, params, states, rng, ep_step, ep_ret, trunc, T, limit)
    D = obs + 15
    got3, want3 = data.reshape(T + 1, n, D), want.reshape(T + 1, n, D)
    assert np.array_equal(env.get_rng(), rng)                                     # every draw (resets, sampling) in the same order
    assert np.array_equal(got3[:T, :, obs + 10], want3[:T, :, obs + 10])           # terminated
    assert np.array_equal(got3[:T, :, obs + 11], want3[:T, :, obs + 11])           # truncated
    assert want3[:T, :, obs + 11].sum() >= n * (T // limit)                        # resets really happened
    close(got3[..., :obs], want3[..., :obs], 2e-3, 2e-4, "observations (incl. the final row)")
    close(got3[:T, :, obs:obs + 8], want3[:T, :, obs:obs + 8], 2e-3, 2e-3, "action means / actions")
    close(got3[:T, :, obs + 8], want3[:T, :, obs + 8], 1e-3, 1e-3, "log-prob")
    close(got3[:T, :, obs + 9], want3[:T, :, obs + 9], 2e-3, 2e-2, "reward")
    assert np.all(got3[..., obs + 12:] == 0)                                      # learner columns: zero-filled by collect (dead until evaluate / GAE write them)
    # the north-star's 1e-4 bound, free of closed-loop amplification (a random actor is chaotic across resets, which is what the 2e-3 above absorbs):
    # (1) before the first reset (step limit 12) the trajectories themselves agree to 1e-4 relative
    K = 8
    close(got3[:K, :, :obs], want3[:K, :, :obs], 1e-4, 3e-5, "observations, first steps")                       # 1e-4 of the quantities' scale (floor 0.3)
    close(got3[:K, :, obs:obs + 8], want3[:K, :, obs:obs + 8], 1e-4, 3e-5, "action means / actions, first steps")
    close(got3[:K, :, obs + 9], want3[:K, :, obs + 9], 1e-4, 1e-4, "reward, first steps")
    # (2) every step re-anchored on the engine's own rows: actor mean from the recorded observation, log-prob from the recorded mean / action
    rows = np.ascontiguousarray(got3[:T].reshape(-1, D))
    pol_mean = port.make_policy(blob[:-4].copy(), arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=B.HEAD_IDENTITY)
    mean_want, _, _ = port.policy_evaluate_step(pol_mean, np.ascontiguousarray(rows[:, :obs]))
    close(rows[:, obs:obs + 4], mean_want, 1e-4, 2e-5, "actor means re-anchored on the recorded observations")
    log_std = blob[-4:].astype(np.float64)
    z = (rows[:, obs + 4:obs + 8].astype(np.float64) - rows[:, obs:obs + 4]) / np.exp(log_std)
    lp_want = (-0.5 * z * z - log_std - 0.5 * np.log(2 * np.pi)).sum(1)
    close(rows[:, obs + 8], lp_want, 1e-4, 1e-4, "log-prob re-anchored on the recorded means / actions")
    close(env.get_parameters(), params, 2e-6, 0, "parameters after the in-kernel resets")
    close(env.get_state(), states, 5e-3, 2e-3, "final states (incl. the action-history ring)")


def test_collect_reset_warp_variant_is_bit_identical(tmp_path):
    """k_collect_lag (B200L2F_COLLECT_LAG=1: resets on a fifth warp, terminated lanes sit out, per-lane step counters; experimental, off by default) writes the same
    dataset, parameters, states and RNG streams as k_collect_ts, bit for bit -- the kernel choice is read once per process, hence two subprocesses"""
    import subprocess
    import sys
    script = tmp_path / "collect_once.py"
    script.write_text("""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import raptor_b200 as rb
from conftest import random_mlp_blob, FOUNDATION_DR
n, T, limit, obs = 333, 48, 11, 22
env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
row = env.get_environment_parameters(); d = FOUNDATION_DR
row[124:139] = np.array([d["t2w"][0], d["t2w"][1], d["t2i"][0], d["t2i"][1], d["mass"][0], d["mass"][1], d["size_dev"], d["tau_rise"][0], d["tau_rise"][1],
                         d["tau_fall"][0], d["tau_fall"][1], d["kq"][0], d["kq"][1], 0.0, d["dist_force"]], np.float32)
env.set_environment_parameters(row)
env.initialize_rng(5, warmup=16); env.initial_parameters(); env.initial_state()
env.load_policy(random_mlp_blob(np.random.RandomState(3), obs, 4, True, True), arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1,
                head=rb.HEAD_PPO_GAUSSIAN, gemm=rb.GEMM_TCGEN05_3XTF32)
env.collect_reset()
a = env.collect(T, limit); b = env.collect(T, limit)          # the second call starts from the carried-over runner state
print(env.last_kernel())
np.savez(sys.argv[1], a=a, b=b, params=env.get_parameters(), state=env.get_state(), rng=env.get_rng())
""" % (ROOT, os.path.join(ROOT, "tests")))
    outs = {}
    for lag in ("0", "1"):
        out = str(tmp_path / ("lag%s.npz" % lag))
        r = subprocess.run([sys.executable, str(script), out], capture_output=True, text=True, timeout=300, env=dict(os.environ, B200L2F_COLLECT_LAG=lag))
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.strip().splitlines()[-1] == ("k_collect_lag" if lag == "1" else "k_collect_ts")
        outs[lag] = np.load(out)
    for k in ("a", "b", "params", "state", "rng"):
        assert np.array_equal(outs["0"][k].view(np.uint32) if outs["0"][k].dtype == np.float32 else outs["0"][k],
                              outs["1"][k].view(np.uint32) if outs["1"][k].dtype == np.float32 else outs["1"][k]), k
    assert outs["0"]["a"][:, 33].sum() > 100                                     # resets happened


def test_ppo_collect_write_back_paths(rb, port):
    """the two write-back paths of k_collect_ts give the same dataset as the oracle: n = 203 (n % 4 != 0: a warp's 32 rows are not 16-byte aligned, every warp takes
    the element loop) and n = 256 into a device buffer whose base is offset by one row (unaligned base: element loop) vs the aligned bulk-copy path on the same inputs"""
    import torch
    spec, obs, D, T, limit = B.SPEC_RAPTOR_DR, 22, 37, 24, 9
    rs = np.random.RandomState(15)
    blob = random_mlp_blob(rs, obs, 4, True, True)
    env_p = foundation_dr_env_params(port, spec)
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=B.HEAD_PPO_GAUSSIAN)

    def fresh(n):
        env = rb.VectorEnvironment(n, spec)
        env.set_environment_parameters(env_p)
        env.initialize_rng(41, warmup=16); env.initial_parameters(); env.initial_state()
        env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN, gemm=rb.GEMM_TCGEN05_3XTF32)
        env.collect_reset()
        return env
    # ---- n % 4 != 0
    n = 203
    env = fresh(n)
    rng, params, states = env.get_rng(), env.get_parameters(), env.get_state()
    got3 = env.collect(T, limit).reshape(T + 1, n, D)
    ep_step = np.zeros(n, np.int32); ep_ret = np.zeros(n, np.float32); trunc = np.ones(n, np.uint8)
    want3 = port.collect(spec, pol, env_p, params, states, rng, ep_step, ep_ret, trunc, T, limit).reshape(T + 1, n, D)
    assert np.array_equal(env.get_rng(), rng)
    assert np.array_equal(got3[:T, :, obs + 10:obs + 12], want3[:T, :, obs + 10:obs + 12]) and want3[:T, :, obs + 11].sum() >= 2 * n
    close(got3[:8, :, :obs + 8], want3[:8, :, :obs + 8], 1e-4, 3e-5, "observations / means / actions, first steps (n = 203)")
    close(got3[..., :obs], want3[..., :obs], 2e-3, 2e-4, "observations (n = 203)")
    assert np.all(got3[..., obs + 12:] == 0) and np.all(got3[T, :, obs:] == 0)
    # ---- aligned (bulk copy) vs unaligned base (element loop): bit-identical datasets
    n = 256
    rows = (T + 1) * n
    a = fresh(n); buf_a = torch.full((rows, D), 3.0, dtype=torch.float32, device="cuda")
    a.collect(T, limit, buf_a); a.synchronize()
    b = fresh(n); big = torch.full((rows + 1, D), 3.0, dtype=torch.float32, device="cuda")
    buf_b = big[1:]                                                        # base + 148 bytes: not a multiple of 16
    assert buf_b.data_ptr() % 16 != 0 and buf_b.is_contiguous()
    b.collect(T, limit, buf_b); b.synchronize()
    assert torch.equal(buf_a, buf_b) and float(big[0].min().item()) == 3.0 and int((buf_a == 3.0).sum().item()) == 0
    assert np.array_equal(a.get_rng(), b.get_rng()) and np.array_equal(a.get_parameters(), b.get_parameters())


@pytest.mark.parametrize("gemm", ["fp32", "tcgen05", "tcgen05-edited-parameters", "tcgen05-device-buffers"])
@pytest.mark.parametrize("spec,sample_parameters", [(B.SPEC_TEACHER, True), (B.SPEC_TEACHER_DR, True), (B.SPEC_TEACHER_DR, False)])
def test_off_policy_steps_vs_oracle(rb, port, spec, sample_parameters, gemm):
    """SAC-teacher collection (rl::components::off_policy_runner step x T) into per-environment replay rings: 200 environments (one full tile + a
    ragged one), 70 steps into 48-row rings (they wrap), step limit 20 + a tightened position threshold (episodes end both ways), a second call
    continues from the carried-over runner state; CUDA-core kernel (k_off_policy) and tensor-core kernel (k_off_policy_ts), host and device rings"""
    import torch
    edited = gemm.endswith("edited-parameters")
    on_device = gemm.endswith("device-buffers")
    g = rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32
    n, T, limit, capacity, obs = 200, 35, 20, 48, 26
    rs = np.random.RandomState(17)
    blob = random_mlp_blob(rs, obs, 8, False, False)
    blob[-4:] += np.float32(-1.0)                   # log_std biases: moderate exploration noise
    env_p = foundation_dr_env_params(port, spec) if spec == B.SPEC_TEACHER_DR else port.nominal_parameters(spec).copy()
    env_p[115] = 0.7                                # termination.position_threshold
    env = rb.VectorEnvironment(n, spec)
    env.set_environment_parameters(env_p)
    env.initialize_rng(43, warmup=16)
    env.initial_parameters()
    env.initial_state()
    env.load_policy(blob, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=8, standardize=0, head=rb.HEAD_SQUASH_EVAL, gemm=g)
    env.collect_reset()
    rng = env.get_rng()
    params, states = env.get_parameters(), env.get_state()
    if edited:   # as in test_ppo_collect_vs_oracle: per-environment reward scale / constant, in effect until the first reset
        params[:, 95] *= rs.uniform(0.5, 1.5, n).astype(np.float32)
        params[:, 96] += rs.uniform(0.0, 0.2, n).astype(np.float32)
        env.set_parameters(params)
    replay = env.new_replay_buffers(capacity, device=on_device)
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_SAMPLE)
    runner = B.new_off_policy_runner(n, capacity, obs)
    D = 2 * obs + 7
    for it in range(2):
        env.off_policy_steps(T, limit, replay, sample_parameters=sample_parameters)
        port.off_policy_steps(spec, pol, env_p, params, states, rng, runner, T, limit, sample_parameters=sample_parameters)
        env.synchronize()                                                              # device buffers: the call only enqueues on the engine's stream
        got = {k: (v.cpu().numpy() if on_device else v) for k, v in replay.items()}
        assert np.array_equal(env.get_rng(), rng), it                                  # every draw (resets, exploration, noise) in the same order
        for k, w in (("position", "position"), ("full", "full"), ("current_episode_start", "current_episode_start"), ("episode_start", "episode_start")):
            assert np.array_equal(got[k], runner[w]), (it, k)
        ep_step, ep_ret, trunc = env.get_runner_state()
        assert np.array_equal(ep_step, runner["episode_step"]) and np.array_equal(trunc, runner["truncated"])
        close(ep_ret, runner["episode_return"], 2e-3, 2e-2, "episode_return")
        gd, wd = got["data"], runner["replay"]
        assert np.array_equal(gd[..., D - 2:], wd[..., D - 2:]), "terminated / truncated flags"
        close(gd[..., :obs], wd[..., :obs], 2e-3, 2e-4, "observations")
        close(gd[..., obs:obs + 4], wd[..., obs:obs + 4], 2e-3, 2e-3, "actions")
        close(gd[..., obs + 4], wd[..., obs + 4], 2e-3, 2e-2, "rewards")
        close(gd[..., obs + 5:2 * obs + 5], wd[..., obs + 5:2 * obs + 5], 2e-3, 2e-4, "next observations")
        if it == 0:   # the north-star's 1e-4 bound before closed-loop amplification sets in: the first ring rows (steps 0..7, before any reset)
            K = 8
            close(gd[:, :K, :obs + 4], wd[:, :K, :obs + 4], 1e-4, 3e-5, "observations / actions, first steps")   # 1e-4 of the quantities' scale (floor 0.3)
            close(gd[:, :K, obs + 4], wd[:, :K, obs + 4], 1e-4, 1e-4, "rewards, first steps")
            close(gd[:, :K, obs + 5:2 * obs + 5], wd[:, :K, obs + 5:2 * obs + 5], 1e-4, 3e-5, "next observations, first steps")
        close(env.get_parameters(), params, 2e-6, 0, "parameters after the in-kernel resets")
        close(env.get_state(), states, 2e-3, 2e-4, "states")
    assert runner["full"].all() and runner["replay"][..., D - 2].sum() > 0 and runner["replay"][..., D - 1].sum() > runner["replay"][..., D - 2].sum()
    # a row's next observation is the following row's observation inside an episode (no observation noise in these specs)
    gd = got["data"]
    e0 = gd[0]
    pos = int(got["position"][0])
    order = np.r_[pos:capacity, 0:pos]              # oldest -> newest
    for a_, b_ in zip(order[:-1], order[1:]):
        if e0[a_, D - 1] == 0:
            assert np.array_equal(e0[a_, obs + 5:2 * obs + 5], e0[b_, :obs])
    # the learner-side read (rl_tools::gather_batch, SEQUENCE_LENGTH 1): a pure gather of the rings, bit-exact vs the oracle on the same rings
    ring_host = dict(replay=got["data"], position=got["position"], full=got["full"])
    for (b0, cnt, B_) in ((0, None, 300), (64, 10, 37)):
        rng_b = port.rng_states(900 + B_, B_, warmup=3)
        want_b = port.gather_batch(ring_host, rng_b.copy(), 500, env_begin=b0, env_count=cnt)
        rng_in = torch.from_numpy(rng_b.view(np.int64).copy()).cuda() if on_device else rng_b.copy()
        got_b = env.gather_batch(replay, rng_in, 500, env_begin=b0, env_count=cnt)
        env.synchronize()
        for k, v in want_b.items():
            g_ = got_b[k].cpu().numpy() if on_device else got_b[k]
            assert np.array_equal(g_, v), k
        if cnt:
            assert want_b["env_index"].min() >= b0 and want_b["env_index"].max() < b0 + cnt
    fresh = env.new_replay_buffers(capacity, device=on_device)
    with pytest.raises(rb.EngineError, match="at least one element"):
        env.gather_batch(fresh, torch.zeros(4, dtype=torch.int64, device="cuda") + 12345 if on_device else np.full(4, 12345, np.uint64))
    with pytest.raises(rb.EngineError, match="SAC actor"):
        env.load_policy(random_mlp_blob(rs, obs, 4, False, False), arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=0, head=rb.HEAD_IDENTITY, gemm=g)
        env.off_policy_steps(1, limit, replay)


@pytest.mark.gpu
@pytest.mark.parametrize("on_device", [False, True])
def test_gather_batch_sequential_vs_oracle(rb, port, on_device):
    """rl_tools::gather_batch for SEQUENCE_LENGTH > 1 (recurrent SAC; operations_generic.h:240-434): the CUDA walk against the oracle (itself bit-exact against the
    reference's gather_batch_step for the same parameter sets, tests/test_oracle_vs_reference.py::test_gather_batch_sequential) -- fixed / random sequence
    lengths, with / without the nominal-length draw, from any row / from the episode's first row, both `next_*` view offsets, rings partially filled and wrapped,
    an environment sub-range; host and device buffers.  Bit-exact: a pure gather plus integer / double RNG arithmetic."""
    import torch
    n, obs, mel = 40, 26, 500
    env = rb.VectorEnvironment(n, B.SPEC_TEACHER)
    cases = [dict(sequence_length=8, include_first_step_in_targets=True, always_sample_from_initial_state=False, random_seq_length=False, capacity=48),
             dict(sequence_length=8, capacity=640),                                                        # the reference's defaults for L > 1: all three switches on
             dict(sequence_length=8, include_first_step_in_targets=False, always_sample_from_initial_state=False, random_seq_length=True,
                  enable_nominal_sequence_length_probability=False, capacity=48),
             dict(sequence_length=24, nominal_sequence_length_probability=0.1, capacity=640),
             dict(sequence_length=2, capacity=640),
             dict(sequence_length=1, capacity=48),                                                         # L = 1 through the general walk (defaults: all switches off)
             dict(sequence_length=130, nominal_sequence_length_probability=0.3, capacity=700)]              # longer than a warp, than most episodes
    for ci, case in enumerate(cases):
        case = dict(case)
        cap = case.pop("capacity")
        for wrapped in (False, True):
            rs = np.random.RandomState(50 * ci + wrapped)
            fill = [cap + 11 * e + 5 for e in range(n)] if wrapped else [mel + 30 + e if cap > mel else cap * 3 // 4 for e in range(n)]
            ring = B.synthetic_replay_rings(rs, n, cap, obs, fill=fill)
            for (b0, cnt, Bsz) in ((0, None, 97), (8, 5, 33)):
                rng_b = port.rng_states(7000 + ci, Bsz, warmup=3)
                rng_want = rng_b.copy()
                kw = dict(case); Lq = kw.pop("sequence_length")
                dflt = Lq > 1
                want = port.gather_batch_sequential(ring, rng_want, mel, Lq, include_first_step_in_targets=kw.get("include_first_step_in_targets", dflt),
                                                    always_sample_from_initial_state=kw.get("always_sample_from_initial_state", dflt),
                                                    random_seq_length=kw.get("random_seq_length", dflt),
                                                    enable_nominal_sequence_length_probability=kw.get("enable_nominal_sequence_length_probability", True),
                                                    nominal_sequence_length_probability=kw.get("nominal_sequence_length_probability", 0.5), env_begin=b0, env_count=cnt)
                replay = dict(data=ring["replay"], episode_start=ring["episode_start"], position=ring["position"], full=ring["full"])
                rng_in = rng_b.copy()
                if on_device:
                    replay = {k: torch.from_numpy(v).cuda() for k, v in replay.items()}
                    rng_in = torch.from_numpy(rng_b.view(np.int64).copy()).cuda()
                got = env.gather_batch(replay, rng_in, mel, env_begin=b0, env_count=cnt, **case)
                env.synchronize()
                g = lambda v: v.cpu().numpy() if on_device else v  # noqa: E731
                assert np.array_equal(g(rng_in).view(np.uint64), rng_want), (case, wrapped)
                for k, v in want.items():
                    assert np.array_equal(g(got[k]), v), (case, wrapped, k)
    with pytest.raises(rb.EngineError, match="capacity >= max_episode_length"):
        ring = B.synthetic_replay_rings(np.random.RandomState(1), n, 48, obs, fill=[30] * n)
        env.gather_batch(dict(data=ring["replay"], episode_start=ring["episode_start"], position=ring["position"], full=ring["full"]), port.rng_states(1, 4), mel, sequence_length=4)
    with pytest.raises(rb.EngineError, match="at least one element"):
        ring = B.synthetic_replay_rings(np.random.RandomState(1), n, 48, obs, fill=[0] * n)
        env.gather_batch(dict(data=ring["replay"], episode_start=ring["episode_start"], position=ring["position"], full=ring["full"]), port.rng_states(1, 4), mel,
                         sequence_length=4, always_sample_from_initial_state=False)


@pytest.mark.gpu
@pytest.mark.parametrize("spec,n,gemm", [(B.SPEC_RAPTOR, 256, "tcgen05"), (B.SPEC_RAPTOR_DR, 200, "tcgen05"), (B.SPEC_TEACHER_DR, 203, "tcgen05"), (B.SPEC_RAPTOR, 203, "fp32"),
                                         (B.SPEC_DEFAULT, 203, "fp32"), (B.SPEC_DEFAULT_DR, 200, "tcgen05")])
def test_learner_feed_vs_oracle(rb, port, spec, n, gemm):
    """critic values, GAE and the running normalizer on the collected dataset (the PPO loop step between collect and train) against the oracle,
    which is pinned bit-for-bit to the reference's own evaluate / estimate_generalized_advantages / running_normalizer update
    (tests/test_oracle_vs_reference.py::test_collect_gae_normalizer).  n = 256: every warp takes the TMA bulk path; 200: a ragged last tile;
    203: row pitch not 16-byte aligned (generic loads)."""
    import torch
    gemm = rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32
    T, limit = 24, 9
    obs = port.observation_dim(spec)
    D = obs + 15
    rs = np.random.RandomState(17)
    actor = random_mlp_blob(rs, obs, 4, True, True)
    critic = random_mlp_blob(rs, obs, 1, True, False)
    critic[-65:] *= 20.0        # values of order 1..10 so that the advantage recursion is exercised with realistic magnitudes
    env = rb.VectorEnvironment(n, spec)
    env_p = foundation_dr_env_params(port, spec) if spec in (B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR, B.SPEC_DEFAULT_DR) else port.nominal_parameters(spec)
    env.set_environment_parameters(env_p)
    env.initialize_rng(3, warmup=16)
    env.initial_parameters()
    env.initial_state()
    env.load_policy(actor, arch=rb.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=rb.HEAD_PPO_GAUSSIAN)
    env.load_critic(critic, standardize=1, gemm=gemm)
    env.collect_reset()
    data = env.collect(T, limit)
    assert data[:T * n, obs + 11].sum() > 0
    crit = port.make_policy(critic, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=1, standardize=1, head=B.HEAD_IDENTITY)
    want = data.copy()
    port.evaluate_values(crit, want, n, T)
    # (1) values, host dataset
    got = env.evaluate_values(data.copy(), T)
    close(got[:, obs + 12], want[:, obs + 12], 1e-4, 1e-4, "critic values")
    assert np.array_equal(got[:, :obs + 12], data[:, :obs + 12]) and np.all(got[:, obs + 13:] == 0)
    # (2) stand-alone GAE on the ORACLE's values: bit-exact (the recursion is pure fp32 add / mul in the reference's order)
    for ignore in (False, True):
        w = want.copy()
        port.estimate_generalized_advantages(w, n, T, 0.99, 0.95, ignore)
        g = env.estimate_generalized_advantages(want.copy(), T, 0.99, 0.95, ignore)
        assert np.array_equal(g, w), "GAE ignore_termination=%s" % ignore
        assert np.abs(w[:T * n, obs + 13]).max() > 0.1
    # (3) fused single pass == values pass followed by the stand-alone GAE (bit for bit), on a device-resident dataset
    dev = torch.from_numpy(data.copy()).cuda()
    env.values_and_advantages(dev, T, 0.99, 0.95, False)
    env.synchronize()
    two_pass = env.estimate_generalized_advantages(got.copy(), T, 0.99, 0.95, False)
    assert np.array_equal(dev.cpu().numpy(), two_pass)
    w = want.copy()
    port.estimate_generalized_advantages(w, n, T, 0.99, 0.95, False)
    close(two_pass[:T * n, obs + 13:], w[:T * n, obs + 13:], 1e-3, 1e-3, "advantages / target values from the GPU critic")
    # (4) running observation normalizer, two updates
    mean_g, std_g, mean_w, std_w = np.zeros(obs, np.float32), np.ones(obs, np.float32), np.zeros(obs, np.float32), np.ones(obs, np.float32)
    age_g = age_w = 0
    for _ in range(2):
        age_g = env.normalizer_update(data, T, mean_g, std_g, age_g)
        age_w = port.normalizer_update(data, n, T, mean_w, std_w, age_w)
    assert age_g == age_w == 2
    close(mean_g, mean_w, 1e-5, 1e-6, "normalizer mean")
    close(std_g, std_w, 1e-5, 1e-6, "normalizer std")


@pytest.mark.parametrize("gemm", ["tcgen05", "fp32"])
def test_dagger_gather_vs_oracle(rb, port, gemm):
    """gather_epoch for all teachers in one call (student rollout -> compaction -> teacher / student observations -> teacher labels) against the
    oracle's add_to_dataset, which is pinned bit-for-bit to the reference's own (tests/test_oracle_vs_reference.py::test_dagger_add_to_dataset).
    37 teachers x 8 episodes = 296 environments (a ragged last CTA in the rollout), 90 steps, tight position threshold so that episodes end early."""
    import torch
    n_teachers, E, T = 37, 8, 90
    n = n_teachers * E
    spec = rb.SPEC_RAPTOR_DR
    rs = np.random.RandomState(41)
    env = rb.VectorEnvironment(n, spec)
    env_p = foundation_dr_env_params(port, spec)
    env.set_environment_parameters(env_p)
    env.initialize_rng(13, warmup=16)
    env.sample_initial_parameters()
    params = env.get_parameters()
    params = np.repeat(params[::E], E, axis=0)                 # the episodes of a teacher share the teacher's dynamics
    params[:, 115] = 0.45 * params[:, 115]                      # termination.position_threshold
    env.set_parameters(np.ascontiguousarray(params))
    env.sample_initial_state()
    env.load_policy()
    teachers = np.stack([random_mlp_blob(rs, 26, 8, False, False) for _ in range(n_teachers)])
    offsets = rs.uniform(-0.05, 0.05, (n_teachers, 3)).astype(np.float32)
    env.load_teachers(teachers, offsets, episodes_per_teacher=E, gemm=rb.GEMM_FP32_CUDA_CORES if gemm == "fp32" else rb.GEMM_TCGEN05_3XTF32)
    states0, rng0 = env.get_state(), env.get_rng()
    h0, g0 = env.get_hidden()
    rec = env.rollout(T, record=("states", "terminated", "returns", "episode_length"))     # what gather records internally (deterministic)
    env.set_state(states0); env.set_rng(rng0); env.set_hidden(h0, g0)
    got = env.dagger_gather(T)
    rows = got["rows"]
    assert np.array_equal(got["episode_length"], rec["episode_length"]) and np.array_equal(got["returns"], rec["returns"])
    assert rows == int(rec["episode_length"].sum()) and 0.2 * n * T < rows < n * T, rows
    want = port.dagger_add_to_dataset(params, rec["states"][:T], rec["terminated"], rng0.copy(), teachers, offsets, E)
    assert want["rows"] == rows
    for k in ("truncated", "reset"):
        assert np.array_equal(got[k][:rows], want[k][:rows]), k
    assert np.array_equal(got["episode_start"], want["episode_start"][:n])
    assert got["truncated"][:rows].sum() == n and (got["truncated"][rows:] == 0).all()
    close(got["input_student"][:rows], want["input_student"][:rows], 1e-5, 1e-6, "student observations")
    close(got["output_target"][:rows], want["output_target"][:rows], 2e-4, 2e-4, "teacher action targets")
    assert np.abs(want["output_target"][:rows]).mean() > 0.05
    # device-resident dataset, same bits
    env.set_state(states0); env.set_rng(rng0); env.set_hidden(h0, g0)
    cap = n * T
    dev = dict(input_student=torch.zeros((cap, 22), device="cuda"), output_target=torch.zeros((cap, 4), device="cuda"), truncated=torch.zeros(cap, dtype=torch.uint8, device="cuda"),
               reset=torch.zeros(cap, dtype=torch.uint8, device="cuda"), episode_start=torch.zeros(n, dtype=torch.int32, device="cuda"))
    got_d = env.dagger_gather(T, out=dev)
    env.synchronize()
    assert got_d["rows"] == rows
    for k in ("input_student", "output_target", "truncated", "reset", "episode_start"):
        assert np.array_equal(got_d[k].cpu().numpy()[:rows if k != "episode_start" else n], got[k][:rows if k != "episode_start" else n]), k


def test_time_chunked_scheduler_is_transparent(rb):
    """the persistent (tile, time-chunk) work queue of the tensor-core kernel must not change results: 1 chunk vs 7 chunks, bit for bit;
    3000 envs = 24 tiles incl. a ragged one, 333 steps (chunks of unequal length), DEFAULT spec keeps its action ring in HBM across chunks"""
    import os
    for spec in (rb.SPEC_RAPTOR, rb.SPEC_DEFAULT):
        res = []
        for chunks in ("1", "7"):
            os.environ["B200L2F_CHUNKS"] = chunks
            try:
                e = rb.VectorEnvironment(3000, spec)
                e.initialize_rng(5, warmup=16)
                e.sample_initial_state()
                e.load_policy(gemm=rb.GEMM_TCGEN05_3XTF32)
                o = e.rollout(333, record=("returns", "episode_length", "actions"))
                res.append((e.get_state(), e.get_hidden(), e.get_rng(), o))
            finally:
                del os.environ["B200L2F_CHUNKS"]
        (s1, (h1, g1), r1, o1), (s7, (h7, g7), r7, o7) = res
        assert np.array_equal(s1, s7) and np.array_equal(h1, h7) and np.array_equal(g1, g7) and np.array_equal(r1, r7)
        for k in ("returns", "episode_length", "actions"):
            assert np.array_equal(o1[k], o7[k]), k


def test_mlp_tensor_core_rollout_properties(rb):
    """k_rollout_mlp_ts at a size with several waves and time chunks: identical results for 1 and 5 chunks (bit for bit), agreement with the
    CUDA-core kernel (fp32 tolerance), for both H = 1 specs and both heads"""
    import os
    rs = np.random.RandomState(3)
    for spec, in_dim, out_dim, head in ((rb.SPEC_TEACHER, 26, 8, rb.HEAD_SQUASH_EVAL), (rb.SPEC_RAPTOR, 22, 4, rb.HEAD_IDENTITY)):
        blob = random_mlp_blob(rs, in_dim, out_dim, out_dim == 4, False)
        res = {}
        for key, gemm, chunks in (("fp32", rb.GEMM_FP32_CUDA_CORES, None), ("ts1", rb.GEMM_TCGEN05_3XTF32, "1"), ("ts5", rb.GEMM_TCGEN05_3XTF32, "5")):
            if chunks:
                os.environ["B200L2F_CHUNKS"] = chunks
            try:
                e = rb.VectorEnvironment(2900, spec)
                e.initialize_rng(9, warmup=16)
                e.sample_initial_state()
                e.load_policy(blob, arch=rb.POLICY_MLP, input_dim=in_dim, hidden_dim=64, output_dim=out_dim, standardize=int(out_dim == 4), head=head, gemm=gemm)
                o = e.rollout(30, record=("returns", "episode_length", "actions", "states"))
                res[key] = (e.get_state(), e.get_rng(), o)
            finally:
                os.environ.pop("B200L2F_CHUNKS", None)
        s1, r1, o1 = res["ts1"]; s5, r5, o5 = res["ts5"]; sf, rf, of = res["fp32"]
        assert np.array_equal(s1, s5) and np.array_equal(r1, r5)
        for k in ("returns", "episode_length", "actions", "states"):
            assert np.array_equal(o1[k], o5[k]), k
        assert np.array_equal(r1, rf)
        close_relative(o1["actions"][:15], of["actions"][:15], 1e-4, {"action": (slice(0, 4), 0.1)}, "actions ts vs fp32")
        close_relative(o1["states"][:16], of["states"][:16], 1e-4, STATE_GROUPS, "states ts vs fp32")


def test_axial_dynamics_specialisation(rb, port):
    """the tensor-core kernels drop the zero products of the rotor / inertia matrices when every vehicle thrusts along body z with diagonal
    inertia (k_param_features bit2) and evaluate the remaining terms on the packed fp32 pipe (dynamics_axial_packed).  (1) the axial form agrees
    with the general form on such vehicles to fp32 rounding (different FMA association, 40 closed-loop steps); (2) a tilted rotor and an
    off-diagonal inertia entry switch the general form on, which must agree with the CUDA-core kernel and the oracle"""
    import os
    n, T = 300, 40
    blob = np.load(os.path.join(G, "raptor_kat.npz"))["blob"]

    def run(params=None, gemm=rb.GEMM_TCGEN05_3XTF32, general=False):
        if general:
            os.environ["B200L2F_DYNAMICS"] = "general"
        try:
            e = rb.VectorEnvironment(n, rb.SPEC_RAPTOR)
            e.initialize_rng(17, warmup=16)
            if params is not None:
                e.set_parameters(params)
            e.sample_initial_state()
            e.load_policy(blob, gemm=gemm)
            s0, r0 = e.get_state(), e.get_rng()
            o = e.rollout(T, record=("states", "actions", "rewards"))
            return e, s0, r0, o
        finally:
            os.environ.pop("B200L2F_DYNAMICS", None)

    _, _, _, oa = run()
    _, _, _, og = run(general=True)
    close_relative(oa["states"], og["states"], 2e-5, STATE_GROUPS, "axial vs general form, states")
    close_relative(oa["actions"], og["actions"], 2e-5, {"action": (slice(0, 4), 0.1)}, "axial vs general form, actions")
    close(oa["rewards"], og["rewards"], 1e-4, 1e-4, "axial vs general form, rewards")
    # non-axial vehicles
    e0 = rb.VectorEnvironment(n, rb.SPEC_RAPTOR)
    params = e0.get_parameters()
    rs = np.random.RandomState(2)
    tilt = rs.normal(0, 0.05, (n, 4, 2)).astype(np.float32)
    d = params[:, 12:24].reshape(n, 4, 3)
    d[:, :, 0] = tilt[:, :, 0]; d[:, :, 1] = tilt[:, :, 1]; d[:, :, 2] = np.sqrt(1 - tilt[:, :, 0] ** 2 - tilt[:, :, 1] ** 2)
    J = params[:, 64:73].reshape(n, 3, 3).astype(np.float64)
    J[:, 0, 1] = J[:, 1, 0] = 0.1 * J[:, 0, 0]
    params[:, 64:73] = J.reshape(n, 9).astype(np.float32)
    params[:, 73:82] = np.linalg.inv(J).reshape(n, 9).astype(np.float32)
    et, s0, r0, ot = run(params)
    _, _, _, of = run(params, gemm=rb.GEMM_FP32_CUDA_CORES)
    pol = port.make_policy(blob)
    want = port.rollout(B.SPEC_RAPTOR, pol, params, s0.copy(), r0.copy(), T)
    assert not np.allclose(ot["states"][10], oa["states"][10], atol=1e-4)       # the perturbation matters
    for got, name in ((ot, "tcgen05"), (of, "fp32")):
        close_relative(got["actions"], want["actions"], 1e-4, {"action": (slice(0, 4), 0.1)}, name + " actions")
        close_relative(got["states"], want["states"], 1e-4, STATE_GROUPS, name + " states")
