#!/usr/bin/env python
"""tests/golden/generate_checkpoints.py -- writes the checkpoint code-export fixtures with the REFERENCE's own save_code
(oracle/_ref, built from /root/reference; run in the build container only):

    tests/golden/checkpoints/teacher_sac.h.gz   Sequential<MLP 26-64-64-8, SampleAndSquash>            (SAC teacher actor)
    tests/golden/checkpoints/ppo_actor.h.gz     Sequential<Standardize, mlp_unconditional_stddev 22-64-64-4>  (PPO actor with log_std)
    tests/golden/checkpoints/blobs.npz          the weights that went in (engine blob order) -- what the reader must give back bit for bit

The files are assembled like rl::loop::steps::checkpoint::save_code does (actor, example input / output, meta).  The Raptor checkpoint itself
(Dense-GRU-Dense) is covered by reading the reference's file where it exists and by the committed tests/golden/raptor_kat.npz."""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import binding as B  # noqa: E402

ref = B.Ref()
rs = np.random.RandomState(20251017)
teacher = (rs.standard_normal(26 * 64 + 64 + 64 * 64 + 64 + 64 * 8 + 8) * 0.2).astype(np.float32)
ppo = np.concatenate([rs.standard_normal(22) * 0.5, 0.5 + rs.random_sample(22), rs.standard_normal(22 * 64 + 64 + 64 * 64 + 64 + 64 * 4 + 4) * 0.2,
                      -0.5 + 0.1 * rs.standard_normal(4)]).astype(np.float32)
out = os.path.join(HERE, "checkpoints")
os.makedirs(out, exist_ok=True)
for name, kind, blob, has_std in (("teacher_sac", 1, teacher, 0), ("ppo_actor", 2, ppo, 1)):
    text = ref.save_code(kind, blob, has_std, name="fixtures/" + name)
    with gzip.GzipFile(os.path.join(out, name + ".h.gz"), "wb", mtime=0) as f:
        f.write(text.encode())
    print(name, len(text), "characters")
np.savez_compressed(os.path.join(out, "blobs.npz"), teacher_sac=teacher, ppo_actor=ppo)
