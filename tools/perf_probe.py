"""Perf probe (GPU box): single-team latency and saturated throughput of the solve."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.solver import DynamoSolver  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, inputs = d["t_gyr"], d["inputs"]
    s = DynamoSolver(device=0)
    info = s.info()
    print(os.environ.get("MAG_B200_LIB", "default lib"), info)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    teams = info["evolve_teams"]
    cases = [(5, 1), (24, 1), (5, teams), (24, 4 * teams)]
    for gid, copies in cases:
        ids = np.full(copies, gid, np.int32)
        sub = np.ascontiguousarray(np.repeat(inputs[:, gid - 1:gid, :], copies, axis=1))
        s.run(ids[:1], t, sub[:, :1], profiles=prof, scalars=scal)
        t0 = time.perf_counter()
        r = s.run(ids, t, sub, profiles=prof, scalars=scal)
        dt = time.perf_counter() - t0
        W = int(r["point_stages"].sum())
        tm = s.last_timing()
        print(f"gal {gid} x{copies}: kernels {tm['ms_kernels']:.1f} ms (profile phase {tm['ms_profile_phase']:.1f}), W={W:.3e}, "
              f"{W/(tm['ms_kernels']*1e-3)/1e9:.2f} G point-stages/s, replays {s.info()['fast_path_replays']}")
    n = 4096
    syn = synthetic_histories(inputs, n, first=196)
    ids = galaxy_ids(n, first=196)
    r = s.run(ids, t, syn, profiles=prof, scalars=scal)
    tm = s.last_timing()
    W = int(r["point_stages"].sum())
    print(f"synthetic x{n}: kernels {tm['ms_kernels']:.1f} ms (profile phase {tm['ms_profile_phase']:.1f}), {n/(tm['ms_kernels']*1e-3):.0f} gal/s, "
          f"{W/(tm['ms_kernels']*1e-3)/1e9:.2f} G point-stages/s, replays {s.info()['fast_path_replays']}, max W {int(r['point_stages'].max()):.3e}")
    s.close()


if __name__ == "__main__":
    main()
