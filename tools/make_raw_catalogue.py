"""Writes a synthetic "Galform-shaped" catalogue (magnetizer_b200/synthetic.py: BASELINE configs 3-5) as the
flat binary file the host driver reads with --raw-input:

    "MAGRAW1\\0", int64 ngal, int64 nz, float64 t[nz], float64 inputs[13][ngal][nz]

    python tools/make_raw_catalogue.py <ngal> <out.raw>

(There is no HDF5 WRITER for input files in this repo -- h5py is not available -- and the reference's input
files come from Galform, python/prepare_input.py:93-150; the driver reads real input files through its own
HDF5 reader.)"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.synthetic import synthetic_histories  # noqa: E402


def main():
    ngal, out = int(sys.argv[1]), sys.argv[2]
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    with open(out, "wb") as f:
        f.write(b"MAGRAW1\0" + struct.pack("<qq", ngal, t.size))
        f.write(np.ascontiguousarray(t, "<f8").tobytes())
        blk = 50000
        # column-major over galaxies: write each input column in turn, generated block by block
        cols = [[] for _ in range(13)]
        for g0 in range(0, ngal, blk):
            part = synthetic_histories(base, min(blk, ngal - g0), first=g0)
            for c in range(13):
                cols[c].append(np.ascontiguousarray(part[c], "<f8"))
        for c in range(13):
            for a in cols[c]:
                f.write(a.tobytes())
    print(f"wrote {out}: {ngal} galaxies x {t.size} snapshots, {os.path.getsize(out) / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
