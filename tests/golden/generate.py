"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref/libl2f_ref.so,
built by oracle/Makefile from /root/reference).  Run in the build container:

    python tests/golden/generate.py

Outputs (all float32 / uint64 / int32 / uint8 numpy arrays):
  tests/golden/raptor_kat.npz        the known-answer test that ships inside the Raptor checkpoint
                                     (checkpoint.h:197-214: example::input / example::output) + the weight blob
  tests/golden/default_8x500.npz     BASELINE config 1: default l2f spec, 8 envs x 500 steps, seed 0, Raptor policy
  tests/golden/raptor_dr_64x100.npz  foundation-policy env (H=1, Langevin) with per-env domain-randomised dynamics
  tests/golden/teacher_16x20.npz     pre-training env (OBS 26) observations / steps under random actions
  tests/golden/raptor_noise_8x50.npz observation + action noise on (exercises Box-Muller draws in observe/step)
  raptor_b200/data/raptor_policy_2084.f32   the Raptor weights in the engine's blob layout (product data file)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import binding as B  # noqa: E402
from conftest import foundation_dr_env_params  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def closed_loop(ref, spec, params, rng, T, keep_states):
    n = params.shape[0]
    states0 = ref.sample_initial_state_n(spec, params, rng)
    rng0 = rng.copy()
    s, r = states0.copy(), rng.copy()
    h = np.tile(ref.policy_initial_hidden(), (n, 1)).astype(np.float32)
    g = np.zeros(n, np.int32)
    out = ref.rollout(spec, params, s, r, T, hidden=h, gru_step=g)
    keep = np.array(sorted(set(keep_states)), np.int32)
    return dict(spec=np.int32(spec), params=params, states0=states0, rng0=rng0, state_steps=keep, states=out["states"][keep],
                observations_first=out["observations"][:4], actions=out["actions"], rewards=out["rewards"], terminated=out["terminated"],
                final_hidden=h, final_gru_step=g, final_rng=r)


def main():
    B.build("ref")
    ref = B.Ref()
    blob = ref.policy_export()
    kin, kout = ref.policy_kat_export()
    mean, mx = ref.policy_kat()
    print("reference KAT: mean |d| = %.3g max |d| = %.3g" % (mean, mx))
    np.savez_compressed(os.path.join(OUT, "raptor_kat.npz"), blob=blob, input=kin, output=kout, h0=ref.policy_initial_hidden())
    os.makedirs(os.path.join(ROOT, "raptor_b200", "data"), exist_ok=True)
    blob.tofile(os.path.join(ROOT, "raptor_b200", "data", "raptor_policy_2084.f32"))

    # config 1: default spec, 8 x 500, seed 0 (streams seeded 0..7, no warm-up: exactly initialize_rng(seed 0))
    spec = B.SPEC_DEFAULT
    n, T = 8, 500
    rng = ref.rng_states(0, n)
    params = np.tile(ref.nominal_parameters(spec), (n, 1))
    np.savez_compressed(os.path.join(OUT, "default_8x500.npz"), **closed_loop(ref, spec, params, rng, T, list(range(0, 101)) + list(range(125, 501, 25))))

    # DR foundation-policy env, 64 x 100, warm streams
    spec = B.SPEC_RAPTOR_DR
    n, T = 64, 100
    rng = ref.rng_states(1000, n, warmup=32)
    rng_before_params = rng.copy()
    env_p = foundation_dr_env_params(ref, spec)
    params = ref.sample_initial_parameters_n(spec, env_p, rng)
    d = closed_loop(ref, spec, params, rng, T, range(0, 101, 5))
    d.update(env_params=env_p, rng_before_params=rng_before_params)
    np.savez_compressed(os.path.join(OUT, "raptor_dr_64x100.npz"), **d)

    # teacher env under random actions (observe/step/reward/terminated, no policy)
    spec = B.SPEC_TEACHER
    n, T = 16, 20
    rs = np.random.RandomState(0)
    rng = ref.rng_states(77, n, warmup=32)
    p = ref.nominal_parameters(spec)
    states, obs, acts, rews, terms = [], [], [], [], []
    s = ref.sample_initial_state_n(spec, np.tile(p, (n, 1)), rng)
    rng0 = rng.copy()
    states.append(s.copy())
    for t in range(T):
        a = rs.uniform(-1, 1, (n, 4)).astype(np.float32)
        o = np.zeros((n, 26), np.float32); nx = np.zeros_like(s); rw = np.zeros(n, np.float32); tm = np.zeros(n, np.uint8)
        for i in range(n):
            r = rng[i:i + 1].copy()
            o[i] = ref.observe(spec, p, s[i], r)
            nx[i], _ = ref.step(spec, p, s[i], a[i], r)
            rw[i] = ref.reward(spec, p, s[i], a[i], nx[i])
            tm[i] = ref.terminated(spec, p, nx[i])
            rng[i] = r[0]
        s = nx
        states.append(s.copy()); obs.append(o); acts.append(a); rews.append(rw); terms.append(tm)
    np.savez_compressed(os.path.join(OUT, "teacher_16x20.npz"), spec=np.int32(spec), params=p, rng0=rng0, states=np.array(states), observations=np.array(obs),
                        actions=np.array(acts), rewards=np.array(rews), terminated=np.array(terms), final_rng=rng)

    # noise on
    spec = B.SPEC_RAPTOR
    n, T = 8, 50
    rng = ref.rng_states(5, n, warmup=32)
    p = ref.nominal_parameters(spec)
    p[108:113] = [0.01, 0.02, 0.03, 0.04, 0.05]
    p[113] = 0.05
    params = np.tile(p, (n, 1))
    np.savez_compressed(os.path.join(OUT, "raptor_noise_8x50.npz"), **closed_loop(ref, spec, params, rng, T, range(0, 51, 5)))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
