import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from magnetizer_b200.solver import DynamoSolver
d = np.load("tests/golden/example_input.npz"); t, base = d["t_gyr"], d["inputs"]
s = DynamoSolver(device=0); n = s.info()["evolve_teams"]
prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
for gid, keep in ((5, 3), (5, 6), (5, 9), (5, 12), (5, 20), (5, 40), (23, 40), (22, 40)):
    inputs = np.ascontiguousarray(np.repeat(base[:, gid - 1:gid, :], n, axis=1))
    first = int(np.argmax(inputs[0, 0] >= 0.5)); inputs[0, :, first + keep:] = -1.0e10
    ids = np.full(n, gid, np.int32)
    r = s.run(ids, t, inputs, profiles=prof, scalars=scal); tm = s.last_timing()
    W = int(r["point_stages"].sum())
    print(f"gal {gid} keep {keep}: kernels {tm['ms_kernels']:.1f} ms profile {tm['ms_profile_phase']:.1f}  W={W:.3e}  {W/(tm['ms_kernels']-tm['ms_profile_phase'])/1e6:.1f} G pt-st/s evolve, replays {s.info()['fast_path_replays']} status {r['status'][0].tobytes()}")
