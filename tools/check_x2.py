#!/usr/bin/env python
"""tools/check_x2.py -- k_rollout_raptor_x2 (two environments per thread) against the oracle and against k_rollout_raptor_ts, then timing of both.
Run on the GPU box:  python tools/check_x2.py [--time]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import raptor_b200 as rb  # noqa: E402
from oracle import binding as B  # noqa: E402
from conftest import foundation_dr_env_params  # noqa: E402


def run(kernel, n, T, seed, record, chunks=None):
    os.environ["B200L2F_X2"] = "1" if kernel == "x2" else "0"
    if chunks:
        os.environ["B200L2F_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("B200L2F_CHUNKS", None)
    port = B.Port()
    env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
    env.set_environment_parameters(foundation_dr_env_params(port, rb.SPEC_RAPTOR_DR))
    env.initialize_rng(seed=seed, warmup=16)
    env.sample_initial_parameters()
    env.sample_initial_state()
    env.load_policy()
    p, s, r = env.get_parameters(), env.get_state(), env.get_rng()
    out = env.rollout(T, record=record)
    h, g = env.get_hidden()
    return dict(params=p, state0=s, rng0=r, out=out, state=env.get_state(), rng=env.get_rng(), hidden=h, gru_step=g)


def main():
    port = B.Port()
    pol = port.make_policy(rb.raptor_policy_blob())
    h0 = rb.raptor_policy_blob()[352 + 16 + 2 * (768 + 48):][:16]
    rec = ("states", "actions", "rewards", "terminated", "returns", "episode_length", "observations")
    ok = True
    for n, T, chunks in [(300, 100, None), (1000, 64, 4), (300, 600, None)]:
        x = run("x2", n, T, 7, rec, chunks)
        t = run("ts", n, T, 7, rec, chunks)
        states, rng = x["state0"].copy(), x["rng0"].copy()
        hid = np.tile(h0, (n, 1)).astype(np.float32); gs = np.zeros(n, np.int32)
        want = port.rollout(rb.SPEC_RAPTOR_DR, pol, x["params"], states, rng, T, hidden=hid, gru_step=gs)
        Tc = min(T, 100)
        for name, o in (("x2", x), ("ts", t)):
            da = np.abs(o["out"]["actions"][:Tc] - want["actions"][:Tc]).max()
            ds = np.abs(o["out"]["states"][:Tc + 1] - want["states"][:Tc + 1]).max()
            dob = np.abs(o["out"]["observations"][:Tc] - want["observations"][:Tc, :, :22]).max()
            dr = np.abs(o["out"]["rewards"][:Tc] - want["rewards"][:Tc]).max()
            term_eq = np.array_equal(o["out"]["terminated"], want["terminated"])
            rng_eq = np.array_equal(o["rng"], rng)
            gs_eq = np.array_equal(o["gru_step"], gs)
            dh = np.abs(o["hidden"] - hid).max()
            fin = np.abs(o["state"] - states).max()
            print("n=%d T=%d chunks=%s %s vs oracle: |d action| %.2e |d state| %.2e |d obs| %.2e |d reward| %.2e terminated %s rng %s gru_step %s |d hidden| %.2e |d final state| %.2e (first %d steps)"
                  % (n, T, chunks, name, da, ds, dob, dr, term_eq, rng_eq, gs_eq, dh, fin, Tc), flush=True)
            if name == "x2" and not (da < 2e-4 and ds < 2e-3 and rng_eq and gs_eq and (term_eq or T > 100)):
                ok = False
        print("   x2 vs ts: |d action| %.2e, |d returns| %.2e, episode_length equal %s" % (np.abs(x["out"]["actions"] - t["out"]["actions"]).max(),
              np.abs(x["out"]["returns"] - t["out"]["returns"]).max(), np.array_equal(x["out"]["episode_length"], t["out"]["episode_length"])), flush=True)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    if "--time" in sys.argv:
        dev = torch.device("cuda", 0)
        for n, T in [(65536, 1000), (1048576, 200)]:
            for kernel, chunks in [("ts", None), ("x2", None), ("x2", 8), ("x2", 16), ("x2", 32), ("ts", None), ("x2", None)]:
                os.environ["B200L2F_X2"] = "1" if kernel == "x2" else "0"
                if chunks:
                    os.environ["B200L2F_CHUNKS"] = str(chunks)
                else:
                    os.environ.pop("B200L2F_CHUNKS", None)
                env = rb.VectorEnvironment(n, rb.SPEC_RAPTOR_DR)
                env.set_environment_parameters(foundation_dr_env_params(port, rb.SPEC_RAPTOR_DR))
                env.initialize_rng(seed=1, warmup=16)
                env.sample_initial_parameters(); env.sample_initial_state(); env.load_policy()
                s0 = torch.from_numpy(env.get_state()).to(dev)
                ret = torch.zeros(n, dtype=torch.float32, device=dev)
                ms = []
                for i in range(6):
                    env.set_state(s0); env.policy_reset()
                    env.synchronize()
                    t0 = time.perf_counter()
                    env.rollout(T, out={"returns": ret})
                    env.synchronize()
                    ms.append(1e3 * (time.perf_counter() - t0))
                best = min(ms[2:])
                print("time n=%d T=%d %s chunks=%s: %.3f ms  -> %.3fe9 env-steps/s (mean return %.3f)" % (n, T, kernel, chunks, best, n * T / best / 1e6, float(ret.mean())), flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
