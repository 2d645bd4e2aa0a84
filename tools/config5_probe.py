"""BASELINE config 5 probe (GPU box): synthetic histories at p_nx_ref = 401 with Courant
factors 0.045; throughput of a small batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from magnetizer_b200.solver import DynamoSolver, default_params  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402

d = np.load("tests/golden/example_input.npz")
t, base = d["t_gyr"], d["inputs"]
p = default_params()
p.p_nx_ref, p.p_courant_v, p.p_courant_eta = 401, 0.045, 0.045
s = DynamoSolver(params=p, device=0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
syn, ids = synthetic_histories(base, n, first=196), galaxy_ids(n, first=196)
prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
s.run(ids[:8], t, np.ascontiguousarray(syn[:, :8]), profiles=prof, scalars=scal)
r = s.run(ids, t, syn, profiles=prof, scalars=scal)
tm = s.last_timing()
W = int(r["point_stages"].sum())
print(f"config 5 (p_nx_ref=401, courant 0.045): {n} galaxies, kernels {tm['ms_kernels']:.0f} ms (profile phase {tm['ms_profile_phase']:.0f}), "
      f"{n / (tm['ms_kernels'] * 1e-3):.0f} gal/s, W/galaxy {W / n:.3e}, {W / (tm['ms_kernels'] - tm['ms_profile_phase']) / 1e6:.1f} G pt-st/s evolve, "
      f"replays {s.info()['fast_path_replays']}, {s.info()}")
