"""Round-2 probe (GPU box): what the two evolve kernels and their split buy.

  python tools/r2_probe.py [case ...]      cases: cfg2 share tput big variants  (default: all but big)

  cfg2     BASELINE config 2 (the 196 example galaxies, one GPU): everything on one-warp teams /
           cost split / everything on four-warp teams
  share    one GPU's share of BASELINE config 3 at N = 8 (12 500 synthetic galaxies): the same three,
           the split at several thresholds, two-stream vs programmatic-dependent launch
  tput     saturated throughput of both kernels on a uniform workload (copies of example galaxy 5
           truncated to 3 snapshots: nx = 107, 3 x 20000 steps)
  big      40 000 synthetic galaxies
  variants build-time variants named in MAG_PROBE_LIBS (comma separated .so paths): bit-identity of
           their results against the default build on 512 synthetic galaxies, and their `share` time
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from magnetizer_b200 import solver as S  # noqa: E402
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories  # noqa: E402

PROF, SCAL = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax",)


def timed(s, ids, t, inp, out=None, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = s.run(ids, t, inp, profiles=PROF, scalars=SCAL, out=out)
        wall = (time.perf_counter() - t0) * 1e3
        tm = s.last_timing()
        if best is None or tm["ms_kernels"] < best[1]["ms_kernels"]:
            best = (r, tm, wall)
    return best


def report(tag, s, r, tm, wall, n):
    W = int(r["point_stages"].sum())
    info = s.info()
    ev = tm["ms_kernels"] - tm["ms_profile_phase"]
    print(f"{tag:46s} kernels {tm['ms_kernels']:8.1f} ms (profiles {tm['ms_profile_phase']:6.1f}, evolve {ev:8.1f}) wall {wall:8.1f} | "
          f"{n / (tm['ms_kernels'] * 1e-3):8.0f} gal/s  {W / (ev * 1e-3) / 1e9:6.1f} G pt-stages/s  heavy {info['heavy_galaxies']:5d}  "
          f"replays {info['fast_path_replays']}", flush=True)


def main():
    cases = sys.argv[1:] or ["cfg2", "share", "tput", "variants"]
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, base = d["t_gyr"], d["inputs"]
    s = S.DynamoSolver(device=0)
    print(s.info(), flush=True)

    def modes(tag, ids, inp, alphas=(0.8,)):
        out = s.alloc_outputs(len(ids), t.size, PROF, SCAL, pinned=True)
        ref = None
        for name, mode, alpha in [("light only", 0, 0.0)] + [(f"split alpha={a}", 1, a) for a in alphas] + [("heavy only", 2, 0.0)]:
            s.set_heavy_policy(mode=mode, alpha=alpha)
            r, tm, wall = timed(s, ids, t, inp, out=out)
            report(f"{tag} {name}", s, r, tm, wall, len(ids))
            snap = {k: np.array(r[k]) for k in ("profiles", "scalars", "nsteps", "point_stages")}
            if ref is None:
                ref = snap
            else:
                same = all(np.array_equal(ref[k], snap[k]) for k in ("nsteps", "point_stages"))
                err = np.abs(ref["profiles"] - snap["profiles"]).max()
                print(f"      vs light only: exact fields {'identical' if same else 'DIFFER'}, max abs profile difference {err:.2e}", flush=True)
        s.set_heavy_policy(mode=1, alpha=1.5)

    if "cfg2" in cases:
        ids = np.arange(1, base.shape[1] + 1, dtype=np.int32)
        modes("config 2 (196 example galaxies)", ids, base, alphas=(1.5,))
    if "share" in cases:
        n = 12500
        syn = synthetic_histories(base, n, first=20000)
        ids = galaxy_ids(n, first=20000)
        modes("config 3 share (12500 synthetic)", ids, syn, alphas=tuple(float(a) for a in os.environ.get("MAG_PROBE_ALPHAS", "1.0,1.5,2.0,2.5").split(",")))
    if "mid" in cases:
        n = 25000
        syn = synthetic_histories(base, n, first=40000)
        ids = galaxy_ids(n, first=40000)
        modes("25000 synthetic", ids, syn, alphas=(1.5,))
    if "tput" in cases:
        for copies in (1, 148, 444, 1184):
            inp = np.ascontiguousarray(np.repeat(base[:, 4:5, :], copies, axis=1))
            first = int(np.argmax(inp[0, 0] >= 0.5))
            inp[0, :, first + 3:] = -1.0e10
            ids = np.full(copies, 5, np.int32)
            for name, mode in (("light", 0), ("heavy", 2)):
                s.set_heavy_policy(mode=mode)
                r, tm, wall = timed(s, ids, t, inp)
                report(f"uniform gal 5 x{copies} (3 snapshots) {name}", s, r, tm, wall, copies)
        s.set_heavy_policy(mode=1, alpha=1.5)
    if "prof" in cases:
        n = 12500
        syn = synthetic_histories(base, n, first=20000)
        ids = galaxy_ids(n, first=20000)
        s.set_heavy_policy(mode=1, alpha=1.5)
        r, tm, wall = timed(s, ids, t, syn, reps=3)
        report("profile phase probe (12500 synthetic)", s, r, tm, wall, n)
    if "part" in cases:
        # the multi-GPU partition on ONE GPU: one or two host threads (contexts) per GPU, chunk sizes
        n = int(os.environ.get("MAG_PROBE_PART_N", "50000"))
        syn = synthetic_histories(base, n, first=100000)
        ids = galaxy_ids(n, first=100000)
        for threads in ("1", "2"):
            os.environ["MAG_PARTITION_THREADS_PER_GPU"] = threads
            part = S.Partition([0])
            del os.environ["MAG_PARTITION_THREADS_PER_GPU"]
            out = part.alloc_outputs(n, t.size, PROF, SCAL, pinned=True)
            for chunk in (6250, 12500, 25000):
                part.solve(ids[:2000], t, np.ascontiguousarray(syn[:, :2000]), chunk=1000, profiles=PROF, scalars=SCAL)
                t0 = time.perf_counter()
                r = part.solve(ids, t, syn, chunk=chunk, profiles=PROF, scalars=SCAL, out=out)
                dt = time.perf_counter() - t0
                st = part.stats()
                print(f"partition 1 GPU x {threads} thread(s), {n} galaxies, chunk {chunk:6d}: {dt * 1e3:8.1f} ms  {n / dt:8.0f} gal/s  "
                      f"kernel-busy {st['busy_ms'][0]:8.1f} ms", flush=True)
            for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"):
                S.host_free(out[k])
            part.close()
    if "big" in cases:
        n = 40000
        syn = synthetic_histories(base, n)
        ids = galaxy_ids(n)
        modes("40000 synthetic", ids, syn, alphas=(0.8,))
    if "variants" in cases and os.environ.get("MAG_PROBE_LIBS"):
        n = 512
        syn = synthetic_histories(base, n, first=30000)
        ids = galaxy_ids(n, first=30000)
        s.set_heavy_policy(mode=0)
        want = s.run(ids, t, syn)
        n2 = int(os.environ.get("MAG_PROBE_VARIANT_N", "12500"))          # 25000 + first 40000: the probe's "mid" chunk (throughput, no tail)
        first2 = int(os.environ.get("MAG_PROBE_VARIANT_FIRST", "20000"))
        syn2 = synthetic_histories(base, n2, first=first2)
        ids2 = galaxy_ids(n2, first=first2)
        r, tm, wall = timed(s, ids2, t, syn2)
        report(f"default build, light only, {n2}", s, r, tm, wall, n2)
        s.close()
        for lib in os.environ["MAG_PROBE_LIBS"].split(","):
            S._LIB = None
            S.LIB_PATH = lib if os.path.isabs(lib) else os.path.join(ROOT, lib)
            v = S.DynamoSolver(device=0)
            v.set_heavy_policy(mode=0)
            got = v.run(ids, t, syn)
            same = all(np.array_equal(np.asarray(want[k]), np.asarray(got[k]), equal_nan=True)
                       for k in ("profiles", "scalars", "nsteps", "point_stages", "flags")) and \
                np.array_equal(want["status"], got["status"])
            print(f"{os.path.basename(lib)}: results {'bit-identical to' if same else 'DIFFER from'} the default build "
                  f"(max abs profile difference {np.abs(want['profiles'] - got['profiles']).max():.2e})", flush=True)
            r, tm, wall = timed(v, ids2, t, syn2)
            report(f"{os.path.basename(lib)} light only, {n2}", v, r, tm, wall, n2)
            v.close()
        return
    s.close()


if __name__ == "__main__":
    main()
