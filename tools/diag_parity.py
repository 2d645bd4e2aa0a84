"""Diagnostic: per-quantity / per-snapshot row-relative error of the CUDA path against the
CPU oracle for a few galaxies of the example input, fast path and generic path.
Usage (on a GPU box): python tools/diag_parity.py [gal_id ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from magnetizer_b200.solver import DynamoSolver  # noqa: E402
from oracle import oracle_lib  # noqa: E402
from parity import _row_err  # noqa: E402


def main():
    ids = np.array([int(a) for a in sys.argv[1:]] or [1, 5, 24, 162], np.int32)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, inputs = d["t_gyr"], d["inputs"]
    sub = np.ascontiguousarray(inputs[:, ids - 1, :])
    s = DynamoSolver(device=0)
    ora = oracle_lib.solve_batch(s.params, ids, t, sub, want_hist=True)
    for fast in (True, False):
        s.set_fast_path(fast)
        gpu = s.run(ids, t, sub)
        print(f"==== fast_path={fast} replays={s.info()['fast_path_replays']}")
        for k in ("nsteps", "error", "flags", "point_stages"):
            print(k, "equal" if np.array_equal(gpu[k], ora[k]) else f"DIFF gpu={gpu[k]} ora={ora[k]}")
        print("status equal", np.array_equal(gpu["status"], ora["status"]))
        for gi, g in enumerate(ids):
            valid = np.nonzero(ora["nsteps_hist"][gi] > 0)[0]
            print(f"-- galaxy {g}: nx {ora['nx_hist'][gi][valid].tolist()} nsteps {ora['nsteps_hist'][gi][valid].tolist()}")
            for qi, name in enumerate(ora["profile_names"]):
                e = _row_err(gpu["profiles"][qi][gi], ora["profiles"][qi][gi])
                if e.max() > 1e-13:
                    print(f"   {name:8s}", " ".join(f"{x:.0e}" for x in e[valid]))
            for qi, name in enumerate(ora["scalar_names"]):
                a, b = gpu["scalars"][qi][gi], ora["scalars"][qi][gi]
                e = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
                if e.max() > 1e-13:
                    print(f"   {name:8s}", " ".join(f"{x:.0e}" for x in e[valid]))
    s.close()


if __name__ == "__main__":
    main()
