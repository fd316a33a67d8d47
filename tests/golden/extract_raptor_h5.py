#!/usr/bin/env python
"""tests/golden/extract_raptor_h5.py -- copies the HDF5 checkpoint of the published Raptor policy out of the reference's tarball
(/root/reference/data/raptor-policy-checkpoint.tar.gz: <run>/checkpoint.h5, written by rl::loop::steps::checkpoint::save through HighFive) to
tests/golden/checkpoints/raptor_checkpoint.h5, byte for byte.  Data, not source; run in the build container only (the GPU box has no reference tree)."""
import os
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
with tarfile.open("/root/reference/data/raptor-policy-checkpoint.tar.gz") as tar:
    member = next(m for m in tar.getmembers() if m.name.endswith("/checkpoint.h5"))
    data = tar.extractfile(member).read()
out = os.path.join(HERE, "checkpoints", "raptor_checkpoint.h5")
with open(out, "wb") as f:
    f.write(data)
print(member.name, "->", out, len(data), "bytes")
