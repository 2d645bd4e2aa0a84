"""Converts the reference's example input (a data fixture, not source) into a compact
.npz that travels to the GPU box, where /root/reference does not exist.

Run in the build container:  python tests/golden/make_example_fixture.py
Source: /root/reference/example/example_input.hdf5 (196 galaxies x 40 snapshots), read
with the repo's own minimal HDF5 reader (magnetizer_b200/hdf5_min.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from magnetizer_b200.hdf5_min import INPUT_COLUMNS, load_magnetizer_input  # noqa: E402

if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/example/example_input.hdf5"
    t, inputs = load_magnetizer_input(src)
    out = os.path.join(HERE, "example_input.npz")
    np.savez_compressed(out, t_gyr=t, inputs=inputs, columns=np.array(INPUT_COLUMNS))
    print("wrote", out, inputs.shape, os.path.getsize(out), "bytes")
