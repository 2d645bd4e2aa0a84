"""Small MIXED workload for ncu captures (GPU box): n synthetic galaxies, every history
truncated to its first k valid snapshots, so that the teams of an SM run different grid
sizes / loop variants at the same time (unlike tools/ncu_workload.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.solver import DynamoSolver  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    nsteps_max = int(sys.argv[3]) if len(sys.argv) > 3 else 0   # > 0: cap the steps per snapshot, so that no single history sets the kernel time
    first = int(sys.argv[4]) if len(sys.argv) > 4 else 196      # first synthetic galaxy id (40000 + 25000 galaxies + k = 40: the probe's chunk)
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    inputs = synthetic_histories(base, n, first=first)
    ids = galaxy_ids(n, first=first)
    for g in range(n):
        valid = np.nonzero(inputs[0, g] >= 0.5)[0]
        if valid.size > k:
            inputs[0, g, valid[k]:] = -1.0e10
    inputs = np.ascontiguousarray(inputs)
    from magnetizer_b200.solver import default_params
    p = default_params()
    if nsteps_max > 0:
        p.p_nsteps_max = nsteps_max
    s = DynamoSolver(params=p, device=0)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    s.run(ids[:8], t, np.ascontiguousarray(inputs[:, :8]), profiles=prof, scalars=scal)
    r = s.run(ids, t, inputs, profiles=prof, scalars=scal)
    tm = s.last_timing()
    W = int(r["point_stages"].sum())
    ev = tm["ms_kernels"] - tm["ms_profile_phase"]
    print(f"n={n} k={k} evolve {ev:.1f} ms W={W:.4e} -> {W / ev / 1e6:.1f} G pt-stages/s replays={s.info()['fast_path_replays']}")
    s.close()
