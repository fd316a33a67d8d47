"""The plain-C restatement (oracle/l2f_oracle.c) must be BIT-IDENTICAL to the unmodified reference
(oracle/_ref/libl2f_ref.so, built from /root/reference by oracle/Makefile) on every function of the hot
path.  Both are built with -ffp-contract=off, so equality is exact (np.array_equal), not approximate.
Runs only where the reference could be compiled (the build container); on the GPU box the committed
golden fixtures (tests/test_oracle_golden.py) pin the port instead."""
import numpy as np
import pytest

from conftest import foundation_dr_env_params
from oracle import binding as B

SPECS = [B.SPEC_DEFAULT, B.SPEC_DEFAULT_DR, B.SPEC_RAPTOR, B.SPEC_TEACHER, B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR]
DR_SPECS = [B.SPEC_DEFAULT_DR, B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR]


def test_sizes(port, ref):
    for s in SPECS:
        assert port.state_dim(s) == ref.state_dim(s)
        assert port.observation_dim(s) == ref.observation_dim(s)
        assert port.action_history_length(s) == ref.action_history_length(s)


def test_rng_stream(port, ref):
    for seed in [0, 1, 7, 123456789, 2**40 + 5]:
        a = np.array([port.rng_init(seed)], np.uint64)
        b = np.array([ref.rng_init(seed)], np.uint64)
        assert a[0] == b[0]
        for i in range(200):
            if i % 3 == 0:
                x, y = port.rng_uniform(a, -2.0, 3.0), ref.rng_uniform(b, -2.0, 3.0)
            elif i % 3 == 1:
                x, y = port.rng_normal(a, 0.5, 2.0), ref.rng_normal(b, 0.5, 2.0)
            else:
                x, y = port.rng_normal(a, 0.5, 0.0), ref.rng_normal(b, 0.5, 0.0)  # std == 0: no draw
            assert x == y and a[0] == b[0]


@pytest.mark.parametrize("spec", SPECS)
def test_nominal_parameters(port, ref, spec):
    assert np.array_equal(port.nominal_parameters(spec), ref.nominal_parameters(spec))


@pytest.mark.parametrize("spec", DR_SPECS)
def test_sample_initial_parameters_dr(port, ref, spec):
    env_p = foundation_dr_env_params(ref, spec)
    for seed in range(64):
        a = np.array([port.rng_init(seed)], np.uint64)
        for _ in range(seed % 5):
            port.rng_uniform(a, 0, 1)
        b = a.copy()
        pa = port.sample_initial_parameters(spec, env_p, a)
        pb = ref.sample_initial_parameters(spec, env_p, b)
        assert a[0] == b[0]
        assert np.array_equal(pa, pb), np.nonzero(pa != pb)


@pytest.mark.parametrize("spec", SPECS)
def test_initial_and_sampled_state(port, ref, spec):
    p = ref.nominal_parameters(spec)
    assert np.array_equal(port.initial_state(spec, p), ref.initial_state(spec, p))
    for seed in range(64):
        a = np.array([port.rng_init(seed * 7919 + 13)], np.uint64)
        for _ in range(20):
            port.rng_uniform(a, 0, 1)
        b = a.copy()
        pp = p.copy()
        pp[121] = 0.01 * (seed % 3)  # random force std
        pp[123] = 1e-5 * (seed % 2)  # random torque std
        sa = port.sample_initial_state(spec, pp, a)
        sb = ref.sample_initial_state(spec, pp, b)
        assert a[0] == b[0]
        assert np.array_equal(sa, sb), np.nonzero(sa != sb)


@pytest.mark.parametrize("spec", SPECS)
@pytest.mark.parametrize("noise", [False, True])
def test_step_observe_reward_terminated(port, ref, spec, noise):
    rs = np.random.RandomState(spec * 2 + int(noise))
    p0 = ref.nominal_parameters(spec)
    if noise:
        p0[108:113] = [0.01, 0.02, 0.03, 0.04, 0.05]
        p0[113] = 0.05
    for trial in range(24):
        a = np.array([port.rng_init(trial + 1000)], np.uint64)
        for _ in range(30):
            port.rng_uniform(a, 0, 1)
        b = a.copy()
        sa = port.sample_initial_state(spec, p0, a)
        sb = sa.copy()
        b[0] = a[0]
        for t in range(40):
            oa, ob = port.observe(spec, p0, sa, a), ref.observe(spec, p0, sb, b)
            assert np.array_equal(oa, ob)
            act = rs.uniform(-1.2, 1.2, 4).astype(np.float32)
            na, dta = port.step(spec, p0, sa, act, a)
            nb, dtb = ref.step(spec, p0, sb, act, b)
            assert dta == dtb and a[0] == b[0]
            assert np.array_equal(na, nb), (t, np.nonzero(na != nb))
            assert port.reward(spec, p0, sa, act, na) == ref.reward(spec, p0, sb, act, nb)
            assert port.terminated(spec, p0, na) == ref.terminated(spec, p0, nb)
            sa, sb = na, nb


def test_policy_kat_and_single_step(port, ref):
    blob = ref.policy_export()
    pol = port.make_policy(blob)
    kin, kout = ref.policy_kat_export()
    mean, mx = ref.policy_kat()
    assert mean < 5e-7 and mx < 2e-6
    # port through the KAT, exact vs the reference's evaluate_step
    for b in range(2):
        h_p = np.tile(ref.policy_initial_hidden(), (1, 1)).astype(np.float32)
        h_r = h_p.copy()
        st_p = np.zeros(1, np.int32)
        st_r = np.zeros(1, np.int32)
        for t in range(500):
            ap, _, _ = port.policy_evaluate_step(pol, kin[t, b:b + 1], h_p, st_p)
            ar = ref.policy_evaluate_step(kin[t, b:b + 1], h_r, st_r)
            assert np.array_equal(ap, ar) and np.array_equal(h_p, h_r) and st_p[0] == st_r[0]
            assert np.abs(ap - kout[t, b]).max() < 2e-6


@pytest.mark.parametrize("spec,T", [(B.SPEC_DEFAULT, 520), (B.SPEC_RAPTOR, 120), (B.SPEC_RAPTOR_DR, 120)])
def test_closed_loop_rollout(port, ref, spec, T):
    n = 16
    blob = ref.policy_export()
    pol = port.make_policy(blob)
    rng = ref.rng_states(0, n, warmup=16)
    if spec == B.SPEC_RAPTOR_DR:
        params = ref.sample_initial_parameters_n(spec, foundation_dr_env_params(ref, spec), rng)
    else:
        params = np.tile(ref.nominal_parameters(spec), (n, 1))
    states = ref.sample_initial_state_n(spec, params, rng)
    s_a, s_b, r_a, r_b = states.copy(), states.copy(), rng.copy(), rng.copy()
    h_a = np.tile(ref.policy_initial_hidden(), (n, 1)).astype(np.float32)
    h_b = h_a.copy()
    g_a, g_b = np.zeros(n, np.int32), np.zeros(n, np.int32)
    oa = port.rollout(spec, pol, params, s_a, r_a, T, hidden=h_a, gru_step=g_a)
    ob = ref.rollout(spec, params, s_b, r_b, T, hidden=h_b, gru_step=g_b)
    for k in ["states", "observations", "actions", "rewards", "terminated"]:
        assert np.array_equal(oa[k], ob[k]), k
    assert np.array_equal(h_a, h_b) and np.array_equal(g_a, g_b) and np.array_equal(r_a, r_b)


# ---- PPO data path: MLP actor / critic, collect (the reference's own per-environment prologue / epilogue), GAE, running normalizer ----------
from conftest import random_mlp_blob  # noqa: E402


@pytest.mark.parametrize("in_dim,out_dim,standardize", [(22, 4, True), (26, 4, False), (22, 1, True), (26, 1, True), (26, 8, False), (82, 4, True), (82, 1, True)])
def test_mlp_forward(port, ref, in_dim, out_dim, standardize):
    rs = np.random.RandomState(in_dim * 10 + out_dim)
    blob = random_mlp_blob(rs, in_dim, out_dim, standardize, False)
    x = rs.normal(0, 1.0, (64, in_dim)).astype(np.float32)
    pol = port.make_policy(blob, arch=B.POLICY_MLP, input_dim=in_dim, hidden_dim=64, output_dim=out_dim, standardize=int(standardize), head=B.HEAD_IDENTITY)
    got, _, _ = port.policy_evaluate_step(pol, x)
    want = ref.mlp_evaluate(blob, in_dim, out_dim, standardize, x)
    assert np.array_equal(got, want)


def _collect_inputs(lib, spec, n, seed):
    env_p = foundation_dr_env_params(lib, spec) if spec in (B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR, B.SPEC_DEFAULT_DR) else lib.nominal_parameters(spec)
    rng = lib.rng_states(seed, n, warmup=16)
    params = np.tile(env_p, (n, 1)).astype(np.float32)
    states = lib.sample_initial_state_n(spec, params, rng)
    return env_p, params, states, rng


@pytest.mark.parametrize("spec", [B.SPEC_RAPTOR, B.SPEC_RAPTOR_DR, B.SPEC_TEACHER_DR, B.SPEC_DEFAULT, B.SPEC_DEFAULT_DR])
def test_collect_gae_normalizer(port, ref, spec):
    n, T, limit = ref.ppo_sizes()
    obs = port.observation_dim(spec)
    rs = np.random.RandomState(3 + spec)
    actor = random_mlp_blob(rs, obs, 4, True, True)
    critic = random_mlp_blob(rs, obs, 1, True, False)
    env_p, params, states, rng = _collect_inputs(ref, spec, n, 11)
    pol = port.make_policy(actor, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=4, standardize=1, head=B.HEAD_PPO_GAUSSIAN)
    crit = port.make_policy(critic, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=1, standardize=1, head=B.HEAD_IDENTITY)
    A = dict(params=params.copy(), states=states.copy(), rng=rng.copy(), step=np.zeros(n, np.int32), ret=np.zeros(n, np.float32), trunc=np.ones(n, np.uint8))
    Bk = {k: v.copy() for k, v in A.items()}
    for it in range(2):   # second pass starts from carried-over runner state
        da = port.collect(spec, pol, env_p, A["params"], A["states"], A["rng"], A["step"], A["ret"], A["trunc"], T, limit)
        db = ref.collect(spec, actor, True, env_p, Bk["params"], Bk["states"], Bk["rng"], Bk["step"], Bk["ret"], Bk["trunc"])
        assert np.array_equal(da, db), "dataset, pass %d" % it
        for k in A:
            assert np.array_equal(A[k], Bk[k]), k
        assert da[:T * n, obs + 11].sum() > 0, "the fixture must contain truncations"
    # critic values (port) vs the reference MLP, then GAE on both
    port.evaluate_values(crit, da, n, T)
    want_v = ref.mlp_evaluate(critic, obs, 1, True, db[:, :obs])
    assert np.array_equal(da[:, obs + 12], want_v[:, 0])
    db[:, obs + 12] = want_v[:, 0]
    gamma, lam = ref.ppo_gamma_lambda()
    for ignore in (False, True):
        ga, gb = da.copy(), db.copy()
        port.estimate_generalized_advantages(ga, n, T, gamma, lam, ignore)
        ref.estimate_generalized_advantages(spec, gb, ignore)
        assert np.array_equal(ga, gb)
        assert np.abs(ga[:T * n, obs + 13]).max() > 0
    # running normalizer, two updates
    mean_a, std_a, mean_b, std_b = np.zeros(obs, np.float32), np.ones(obs, np.float32), np.zeros(obs, np.float32), np.ones(obs, np.float32)
    age_a = age_b = 0
    for _ in range(2):
        age_a = port.normalizer_update(da, n, T, mean_a, std_a, age_a)
        age_b = ref.normalizer_update(spec, db, mean_b, std_b, age_b)
    assert age_a == age_b == 2 and np.array_equal(mean_a, mean_b) and np.array_equal(std_a, std_b)


def test_dagger_add_to_dataset(port, ref):
    """the port's DAgger data path against the reference's own add_to_dataset (post_training/helper.h:43-110) on a recorded student rollout of
    one teacher: 10 episodes x 500 steps, tight termination thresholds so that episodes end early and the compaction is exercised"""
    ne, T = ref.dagger_sizes()
    spec = B.SPEC_RAPTOR
    rs = np.random.RandomState(23)
    row = ref.nominal_parameters(spec).copy()
    row[115] = 0.6          # termination.position_threshold: some of the sampled initial positions (|p| <= 0.5) drift out
    params = np.tile(row, (ne, 1)).astype(np.float32)
    rng = ref.rng_states(5, ne, warmup=16)
    states = ref.sample_initial_state_n(spec, params, rng)
    pol = port.make_policy(ref.policy_export())
    h = np.tile(ref.policy_initial_hidden(), (ne, 1)).astype(np.float32)
    out = port.rollout(spec, pol, params, states, rng, T, hidden=h, gru_step=np.zeros(ne, np.int32))
    term = out["terminated"].copy()
    term[137:, 3] = 1       # and one forced mid-episode termination
    first = np.where(term.any(0), term.argmax(0), T)
    assert (first < T).sum() >= 1 and (first == T).sum() >= 1, first
    teacher = random_mlp_blob(rs, 26, 8, False, False)
    offset = np.array([0.01, -0.02, 0.03], np.float32)
    got = port.dagger_add_to_dataset(params, out["states"][:T], term, rng.copy(), teacher[None], offset[None], ne)
    want = ref.dagger_add_to_dataset(params, np.ascontiguousarray(out["states"][:T].transpose(1, 0, 2)), np.ascontiguousarray(term.T), teacher, offset)
    rows = want["rows"]
    assert got["rows"] == rows == int(np.minimum(first + 1, T).sum())
    for k in ("input_student", "output_target", "truncated", "reset"):
        assert np.array_equal(got[k][:rows], want[k][:rows]), k
    assert np.array_equal(got["episode_start"][:ne], want["episode_start"][:ne]) and want["episode_start"][1] == first[0] + 1
    assert want["reset"][:rows].all() and want["truncated"][:rows].sum() == ne


# ---- off-policy runner (SAC teacher data collection): the reference's own prologue_per_env / epilogue_per_env / replay-buffer add ------------
@pytest.mark.parametrize("spec,sample_parameters", [(B.SPEC_TEACHER, True), (B.SPEC_TEACHER_DR, True), (B.SPEC_TEACHER_DR, False)])
def test_off_policy_runner_steps(port, ref, spec, sample_parameters):
    """60 runner steps of 16 environments into 48-row replay rings (the ring wraps), step limit 20 plus a tightened position threshold so that
    episodes end both ways; a second call continues from the carried-over runner state.  Rows, ring bookkeeping, parameters, states and RNG
    streams must be identical to the reference's."""
    n, T, limit, capacity = ref.off_policy_sizes()
    obs = port.observation_dim(spec)
    rs = np.random.RandomState(41 + spec)
    actor = random_mlp_blob(rs, obs, 8, False, False)
    actor[-8 + 4:] += np.float32(-1.0)          # log_std biases: moderate exploration noise
    env_p, params, states, rng = _collect_inputs(ref, spec, n, 13)
    env_p = env_p.copy(); env_p[115] = 0.7       # termination.position_threshold
    params[:, 115] = 0.7
    pol = port.make_policy(actor, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_SAMPLE)
    A = dict(params=params.copy(), states=states.copy(), rng=rng.copy(), runner=B.new_off_policy_runner(n, capacity, obs))
    Bk = dict(params=params.copy(), states=states.copy(), rng=rng.copy(), runner=B.new_off_policy_runner(n, capacity, obs))
    for it in range(2):
        port.off_policy_steps(spec, pol, env_p, A["params"], A["states"], A["rng"], A["runner"], T, limit, sample_parameters=sample_parameters, with_states=True)
        ref.off_policy_steps(spec, actor, env_p, Bk["params"], Bk["states"], Bk["rng"], Bk["runner"], sample_parameters=sample_parameters, with_states=True)
        for k in ("params", "states", "rng"):
            assert np.array_equal(A[k], Bk[k]), (it, k)
        for k, v in Bk["runner"].items():
            assert np.array_equal(A["runner"][k], v), (it, k)
    r = Bk["runner"]
    D = 2 * obs + 7
    assert r["full"].all() and (r["position"] == (2 * T) % capacity).all()
    term, trunc = r["replay"][:, :, D - 2], r["replay"][:, :, D - 1]
    assert term.sum() > 0 and (trunc.sum() > term.sum()), "episodes must end by termination and by the step limit"
    assert (np.abs(r["replay"][:, :, obs:obs + 4]) <= 1).all()


def test_gather_batch(port, ref):
    """the learner-side read of the replay rings: SEQUENCE_LENGTH-1 batches (the MLP SAC configuration) against the reference's own gather_batch_step
    on its own SequentialBatch, from partially filled rings (position 35 < capacity, not full) and from wrapped ones (full)"""
    n, T, limit, capacity = ref.off_policy_sizes()
    spec, obs = B.SPEC_TEACHER, 26
    rs = np.random.RandomState(77)
    actor = random_mlp_blob(rs, obs, 8, False, False)
    env_p, params, states, rng = _collect_inputs(ref, spec, n, 19)
    pol = port.make_policy(actor, arch=B.POLICY_MLP, input_dim=obs, hidden_dim=64, output_dim=8, standardize=0, head=B.HEAD_SQUASH_SAMPLE)
    runner = B.new_off_policy_runner(n, capacity, obs)
    Bsz = ref.lib.ref_gather_batch_size()
    mel = ref.lib.ref_gather_batch_max_episode_length()
    assert mel == 500
    for steps in (35, 40):                       # 35 rows per ring, then 75 (wrapped: full)
        port.off_policy_steps(spec, pol, env_p, params, states, rng, runner, steps, limit)
        r_a = port.rng_states(1000 + steps, Bsz, warmup=3)
        r_b = r_a.copy()
        got = port.gather_batch(runner, r_a, mel)
        want = ref.gather_batch(runner, r_b)
        assert np.array_equal(r_a, r_b)
        for k, v in want.items():
            assert np.array_equal(got[k], v), (steps, k)
        assert len(set(got["env_index"].tolist())) > n // 2 and len(set(got["sample_index"].tolist())) > 10
    assert runner["full"].all()


def test_gather_batch_sequential(port, ref):
    """recurrent-SAC batches (SEQUENCE_LENGTH > 1): the restatement against the reference's own gather_batch_step for every compiled parameter set --
    fixed / random sequence lengths, with and without the nominal-length draw, from any row / from the episode's first row, next-views at offset 0 / 1 --
    on rings that are partially filled and on wrapped ones.  Bit-exact, RNG streams included."""
    n = ref.off_policy_sizes()[0]
    obs = 26
    Bsz = ref.lib.ref_gather_batch_size()
    mel = ref.lib.ref_gather_batch_max_episode_length()
    configs = ref.gather_batch_sequential_configs()
    assert len(configs) >= 6 and {c["sequence_length"] for c in configs} >= {2, 8, 24}
    for cfg in configs:
        cap = cfg["capacity"]
        for case, fill in enumerate(([cap * 3 // 4] * n, [cap + 7 * e + 3 for e in range(n)])):   # not full (position > MAX_EPISODE_LENGTH where needed) / wrapped
            rs = np.random.RandomState(100 * cfg["config"] + case)
            runner = B.synthetic_replay_rings(rs, n, cap, obs, fill=fill)
            assert bool(runner["full"].all()) == (case == 1)
            r_a = port.rng_states(4000 + cfg["config"], Bsz, warmup=3)
            r_b = r_a.copy()
            kw = {k: v for k, v in cfg.items() if k not in ("config", "capacity", "sequence_length")}
            got = port.gather_batch_sequential(runner, r_a, mel, cfg["sequence_length"], **kw)
            want = ref.gather_batch_sequential(cfg["config"], runner, r_b)
            assert np.array_equal(r_a, r_b), cfg
            for k, v in want.items():
                assert np.array_equal(got[k], v), (cfg, case, k)
            # the structure the learner relies on: every sample starts with a reset, the last real step of every sequence is a final step
            assert got["reset"][0].all() and got["next_reset"][0].all()
            assert got["final_step_mask"].sum() >= Bsz
            if cfg["random_seq_length"]:
                assert got["final_step_mask"][: cfg["sequence_length"] - 1].sum() > 0   # sequences really end early
