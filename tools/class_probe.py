"""Throughput of the evolve kernel per grid-size class (GPU box): copies of one example galaxy
truncated to its first k valid snapshots, one per team; the difference between k and k-1
isolates one snapshot (its nx decides which fast-loop variant runs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200.solver import DynamoSolver  # noqa: E402

# (galaxy id, snapshots kept, nx of the last kept snapshot in the example input)
CASES = [(5, 2, 107), (5, 3, 107), (30, 2, 107), (30, 3, 136), (85, 1, 107), (85, 2, 172), (23, 1, 107), (23, 2, 187),
         (67, 1, 107), (67, 2, 217), (5, 7, 107), (5, 8, 327)]


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    s = DynamoSolver(device=0)
    info = s.info()
    n = info["evolve_teams"] * int(sys.argv[1]) if len(sys.argv) > 1 else info["evolve_teams"]
    print(info)
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    prev = None
    cases = [c for c in CASES if not os.environ.get("CLASS_PROBE_QUICK") or c[:2] in ((5, 2), (5, 3), (30, 2), (30, 3), (23, 1), (23, 2))]
    for gid, k, nxl in cases:
        inputs = np.ascontiguousarray(np.repeat(base[:, gid - 1:gid, :], n, axis=1))
        first = int(np.argmax(inputs[0, 0] >= 0.5))
        inputs[0, :, first + k:] = -1.0e10
        ids = np.full(n, gid, np.int32)
        s.run(ids[:8], t, np.ascontiguousarray(inputs[:, :8]), profiles=prof, scalars=scal)
        r = s.run(ids, t, inputs, profiles=prof, scalars=scal)
        tm = s.last_timing()
        W = int(r["point_stages"].sum())
        ev = tm["ms_kernels"] - tm["ms_profile_phase"]
        line = f"gal {gid} k={k} (last nx {nxl}): evolve {ev:.1f} ms W={W:.3e} -> {W / ev / 1e6:.1f} G pt-stages/s"
        if prev and prev[0] == gid and prev[1] == k - 1:
            dW, dt = W - prev[2], ev - prev[3]
            line += f" | last snapshot alone: {dW / dt / 1e6:.1f} G pt-stages/s"
        print(line, f"replays {s.info()['fast_path_replays']}", flush=True)
        prev = (gid, k, W, ev)
    s.close()


if __name__ == "__main__":
    main()
