"""tests/golden/extract_dynamics_parameters.py -- copies a few of the reference's 1000 dynamics-parameter JSON files (the wire format its
`json(device, env, parameters)` writes in double precision, rl/environments/l2f/operations_cpu.h:139-411, produced by
src/foundation_policy/pre_training/sample_dynamics_parameters.cpp) out of /root/reference/data/foundation-policy-v1-data.tar.gz.part_a* (a
zstd-compressed tar, read through pyarrow's codec) into tests/golden/dynamics_parameters/.  They are the golden INPUTS of the parameter-JSON
import tests.  Run in the build container only (the GPU box has no /root/reference)."""
import glob
import io
import os
import tarfile

import pyarrow as pa

HERE = os.path.dirname(os.path.abspath(__file__))
IDS = {"0", "1", "181", "500", "646", "863", "873", "999"}


class Cat(io.RawIOBase):
    def __init__(self, paths):
        self.files = [open(p, "rb") for p in paths]
        self.i = 0

    def readable(self):
        return True

    def readinto(self, b):
        while self.i < len(self.files):
            n = self.files[self.i].readinto(b)
            if n:
                return n
            self.i += 1
        return 0


def main():
    parts = sorted(glob.glob("/root/reference/data/foundation-policy-v1-data.tar.gz.part_a*"))
    out = os.path.join(HERE, "dynamics_parameters")
    os.makedirs(out, exist_ok=True)
    done = 0
    z = pa.CompressedInputStream(pa.PythonFile(io.BufferedReader(Cat(parts), 1 << 20), mode="r"), "zstd")
    tf = tarfile.open(fileobj=z, mode="r|")
    for m in tf:
        d, f = os.path.split(m.name)
        if d.startswith("src/foundation_policy/dynamics_parameters_") and f.endswith(".json") and f[:-5] in IDS:
            with open(os.path.join(out, f), "wb") as o:
                o.write(tf.extractfile(m).read())
            done += 1
            if done == len(IDS):
                break
    print("extracted", done)


if __name__ == "__main__":
    main()
