"""Phase clocks of the evolve kernel (GPU box; debug build with -DMAG_PHASE_CLOCKS:
`make -C magnetizer_b200/csrc clk` -> magnetizer_b200/libmag_clk.so): where lane 0 of the
evolve teams spends its cycles."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MAG_B200_LIB"] = os.path.join(ROOT, "magnetizer_b200", "libmag_clk.so")
from magnetizer_b200 import solver as S  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402

NAMES = ["0 record/adjust/seed", "1 time loop total", "2 Bzmod+outputs", "3 prepare_fast_coefs", "4 guard_constants",
         "5 fold_boundary", "6 coef/state load", "7 step loop", "8 sign refresh (in 7)", "9 #refreshes", "10 #loops", "11 #steps"]


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    s = S.DynamoSolver(device=0)
    lib = S.load_library()
    buf = (C.c_ulonglong * 16)()
    prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)
    n = s.info()["evolve_teams"]
    cases = [("gal 24 x teams", np.repeat(base[:, 23:24, :], n, axis=1), np.full(n, 24, np.int32)),
             ("synthetic 4*teams", synthetic_histories(base, 4 * n, first=196), galaxy_ids(4 * n, first=196))]
    for name, inp, ids in cases:
        inp = np.ascontiguousarray(inp)
        s.run(ids[:8], t, np.ascontiguousarray(inp[:, :8]), profiles=prof, scalars=scal)
        lib.mag_debug_phase_clocks(None, 1)
        r = s.run(ids, t, inp, profiles=prof, scalars=scal)
        lib.mag_debug_phase_clocks(buf, 0)
        tm = s.last_timing()
        ev = tm["ms_kernels"] - tm["ms_profile_phase"]
        W = int(r["point_stages"].sum())
        print(f"== {name}: evolve {ev:.1f} ms, W {W:.3e}, {W / ev / 1e6:.1f} G pt-stages/s, teams {n}")
        tot = buf[0] + buf[1] + buf[2]
        print(f"   team-cycles accounted {tot:.3e} = {tot / n / 1.965e6:.1f} ms per team")
        for i, nm in enumerate(NAMES):
            v = buf[i]
            print(f"   {nm:28s} {v:16d}" + (f"  {100 * v / tot:5.1f} %" if i < 9 else ""))
    s.close()


if __name__ == "__main__":
    main()
