"""Perf probe (GPU box): evolve-kernel throughput on pure nx=107 work (example galaxy 5
truncated to its first 6 snapshots = 6 x 20000 steps), one galaxy per team."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
from magnetizer_b200.solver import DynamoSolver  # noqa: E402

d = np.load("tests/golden/example_input.npz")
t, base = d["t_gyr"], d["inputs"]
s = DynamoSolver(device=0)
n = s.info()["evolve_teams"]
prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
for gid, keep, copies in ((5, 6, n), (5, 6, 1), (23, 40, n)):
    inputs = np.ascontiguousarray(np.repeat(base[:, gid - 1:gid, :], copies, axis=1))
    first = int(np.argmax(inputs[0, 0] >= 0.5))
    inputs[0, :, first + keep:] = -1.0e10
    ids = np.full(copies, gid, np.int32)
    s.run(ids[:1], t, inputs[:, :1], profiles=prof, scalars=scal)
    r = s.run(ids, t, inputs, profiles=prof, scalars=scal)
    tm = s.last_timing()
    W = int(r["point_stages"].sum())
    ev = tm["ms_kernels"] - tm["ms_profile_phase"]
    print(f"{os.environ.get('MAG_B200_LIB', 'default')[-14:]} gal {gid} keep {keep} x{copies}: evolve {ev:.1f} ms, {W / ev / 1e6:.1f} G pt-st/s, "
          f"{ev * 1e3 / (W / copies / 107):.3f} us per stage and team, teams/SM {s.info()['evolve_teams_per_sm']}")
