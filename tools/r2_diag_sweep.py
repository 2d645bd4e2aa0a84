"""Diagnostic (GPU box): which evolve path deviates from the oracle for the sweep settings that fail.
For each setting: the 196 example + 32 synthetic galaxies of tests/test_gpu_parity.py::test_parameter_sweep
through the one-warp kernel, the four-warp kernel and the any-nx generic path; per galaxy whose
status / nsteps / point_stages / flags differ from the oracle, the values of all four."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from magnetizer_b200.solver import DynamoSolver  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402
from oracle import oracle_lib  # noqa: E402
from parity import _row_err  # noqa: E402

CASES = {"no_Pdm": dict(p_use_Pdm=0), "ignore_satellite_haloes": dict(p_ignore_satellite_DM_haloes=1)}


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, inputs = d["t_gyr"], d["inputs"]
    syn = synthetic_histories(inputs, 32, first=5000)
    ids = np.concatenate([np.arange(1, inputs.shape[1] + 1, dtype=np.int32), galaxy_ids(32, first=5000)])
    allin = np.ascontiguousarray(np.concatenate([inputs, syn], axis=1))
    prof, scal = ("Br", "Bp", "alp_m", "h"), ("Bmax", "dt")
    for name in sys.argv[1:] or list(CASES):
        p = oracle_lib.default_params()
        for k, v in CASES[name].items():
            setattr(p, k, v)
        ora = oracle_lib.solve_batch(p, ids, t, allin, profiles=prof, scalars=scal)
        s = DynamoSolver(params=p, device=0)
        runs = {}
        for tag, mode, fast in (("light", 0, True), ("heavy", 2, True), ("generic", 0, False)):
            s.set_heavy_policy(mode=mode)
            s.set_fast_path(fast)
            runs[tag] = s.run(ids, t, allin, profiles=prof, scalars=scal)
            print(f"[{name}] {tag}: replays {s.info()['fast_path_replays']}", flush=True)
        s.close()
        for tag, r in runs.items():
            bad = [g for g in range(len(ids)) if r["status"][g].tobytes() != ora["status"][g].tobytes() or r["nsteps"][g] != ora["nsteps"][g]
                   or r["point_stages"][g] != ora["point_stages"][g] or r["flags"][g] != ora["flags"][g]]
            err = max(float(_row_err(r["profiles"][q], ora["profiles"][q]).max()) for q in range(len(prof)))
            print(f"[{name}] {tag}: {len(bad)} galaxies differ in exact fields; worst row-relative profile error {err:.2e}")
            for g in bad[:6]:
                print(f"    idx {g} id {ids[g]}: status {r['status'][g].tobytes().decode()} vs {ora['status'][g].tobytes().decode()}")
                print(f"      nsteps {r['nsteps'][g]} vs {ora['nsteps'][g]}  point_stages {r['point_stages'][g]} vs {ora['point_stages'][g]}  "
                      f"flags {r['flags'][g]} vs {ora['flags'][g]}  dt(last) {r['scalars'][1][g][-1]!r} vs {ora['scalars'][1][g][-1]!r}")


if __name__ == "__main__":
    main()
