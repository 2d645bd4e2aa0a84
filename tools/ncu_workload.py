"""Small fixed workload for ncu captures (GPU box): N copies of the example galaxies
1..196 cycled, one warm-up launch + one measured launch of k_dynamo_run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.solver import DynamoSolver  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    # n copies of ONE medium example galaxy (id 24: 34 snapshots, 55k RK stages, nx 107-191):
    # a bounded kernel (~0.1-0.3 s) whatever n, so that ncu's ~40 replays stay short
    gid = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    inputs = np.ascontiguousarray(np.repeat(base[:, gid - 1:gid, :], n, axis=1))
    if len(sys.argv) > 3:  # truncate the history after this many valid snapshots
        first = int(np.argmax(inputs[0, 0] >= 0.5))
        inputs[0, :, first + int(sys.argv[3]):] = -1.0e10
    ids = np.full(n, gid, np.int32)
    s = DynamoSolver(device=0)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    s.run(ids[:8], t, np.ascontiguousarray(inputs[:, :8]), profiles=prof, scalars=scal)
    r = s.run(ids, t, inputs, profiles=prof, scalars=scal)
    tm = s.last_timing()
    print(f"n={n} kernels {tm['ms_kernels']:.1f} ms W={int(r['point_stages'].sum()):.4e} replays={s.info()['fast_path_replays']}")
    s.close()
