"""Parity sweep (GPU box): a slice of the synthetic catalogue, CUDA path vs CPU oracle, all
21 profiles and 7 scalars.  Usage: python tools/parity_sweep.py [first] [n]"""
import collections
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from magnetizer_b200.solver import DynamoSolver  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402
from oracle import oracle_lib  # noqa: E402
from parity import compare  # noqa: E402

if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    syn, ids = synthetic_histories(base, n, first=first), galaxy_ids(n, first=first)
    s = DynamoSolver(device=0)
    t0 = time.time()
    gpu = s.run(ids, t, syn)
    t1 = time.time()
    ora = oracle_lib.solve_batch(s.params, ids, t, syn)
    t2 = time.time()
    rep = compare(gpu, ora)
    st = collections.Counter(ora["status"].tobytes().decode())
    print(f"synthetic galaxies {first}..{first + n - 1}: GPU {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s, statuses {dict(st)}, "
          f"flags {dict(collections.Counter(ora['flags'].tolist()))}, replays {s.info()['fast_path_replays']}")
    print("worst row-relative error per quantity:", {k: f"{v:.1e}" for k, v in rep.items() if v > 0})
    print("PARITY OK")
