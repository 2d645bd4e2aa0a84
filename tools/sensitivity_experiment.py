"""CPU experiment (TEST INFRASTRUCTURE, runs here): how far do the reference's own results move
when the libraries it leaves unpinned change in their last bits?

The oracle (oracle/magnetizer_oracle.cpp) is built three ways from the same source:

  pinned    -O2 -ffp-contract=off, own fdlibm-style log/exp/pow and Bessel series (oracle/det_math.hpp)
            -- the build every parity test uses
  libm      -DORACLE_LIBM: glibc log/exp/pow, libstdc++ std::cyl_bessel_i/k (what a Fortran build gets
            from libm + GSL up to their own last-bit differences)
  fma       -ffp-contract=fast -march=native: the compiler contracts a*b+c into FMAs (what gfortran -O2
            does by default on an FMA machine: GCC's default is -ffp-contract=fast)
  libm+fma  both

For each variant: Bmax(z=0) of the tutorial galaxies 22, 23, 24 (reference
doc/magnetizer_tutorial.ipynb:673) and, over all 196 example galaxies, the distribution of the
row-relative differences against the pinned build (the parity metric of tests/parity.py).  The
published tutorial values sit inside the band these variants span iff "residual = library noise"
 holds.

    python tools/sensitivity_experiment.py > profiles/r2_sensitivity.md
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle_lib  # noqa: E402

TUTORIAL = {22: 0.45010823, 23: 0.42917456, 24: 0.03083575}
VARIANTS = [
    ("pinned", ["-O2", "-ffp-contract=off"]),
    ("libm", ["-O2", "-ffp-contract=off", "-DORACLE_LIBM"]),
    ("fma", ["-O2", "-ffp-contract=fast", "-march=native"]),
    ("libm+fma", ["-O2", "-ffp-contract=fast", "-march=native", "-DORACLE_LIBM"]),
]
QUANT = ("Br", "Bp", "alp_m", "Bzmod", "h", "n", "Omega", "Shear", "Omega_h", "alp")
SCAL = ("Bmax", "Bavg")


def build(flags, out):
    src = os.path.join(ROOT, "oracle", "magnetizer_oracle.cpp")
    subprocess.check_call(["g++", *flags, "-std=c++17", "-fPIC", "-pthread", "-shared", "-o", out, src])


def run(so, ids, t, inputs):
    oracle_lib._LIB = None
    L = C.CDLL(so)
    L.oracle_solve_batch.restype = C.c_int
    L.oracle_bessel.restype = C.c_double
    L.oracle_bessel.argtypes = [C.c_int32, C.c_double]
    L.oracle_constant.restype = C.c_double
    L.oracle_constant.argtypes = [C.c_int32]
    oracle_lib._LIB = L
    p = oracle_lib.default_params()
    return oracle_lib.solve_batch(p, ids, t, inputs, profiles=QUANT, scalars=SCAL, nthreads=os.cpu_count() or 1)


def row_err(a, b):
    d = np.abs(a - b).max(axis=-1)
    s = np.abs(b).max(axis=-1)
    out = np.zeros_like(d)
    nz = d > 0
    out[nz] = d[nz] / np.where(s[nz] > 0, s[nz], 1e-300)
    return out


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    t, inputs = d["t_gyr"], d["inputs"]
    ids = np.arange(1, inputs.shape[1] + 1, dtype=np.int32)
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, flags in VARIANTS:
            so = os.path.join(tmp, f"liboracle_{name.replace('+', '_')}.so")
            build(flags, so)
            res[name] = run(so, ids, t, inputs)
    base = res["pinned"]
    print("# Sensitivity of the reference's results to its unpinned libraries (round 2)\n")
    print("`python tools/sensitivity_experiment.py` (CPU, this container; g++ " +
          subprocess.check_output(["g++", "-dumpfullversion"], text=True).strip() + ").  See the tool's docstring for the variants.\n")
    print("## Bmax(z=0) of the tutorial galaxies, relative to the value the reference publishes\n")
    print("| build | gal 22 (0.45010823) | gal 23 (0.42917456) | gal 24 (0.03083575) |")
    print("|---|---|---|---|")
    ib = SCAL.index("Bmax")
    for name, _ in VARIANTS:
        r = res[name]
        cells = []
        for g, want in TUTORIAL.items():
            got = r["scalars"][ib][g - 1][-1]
            cells.append(f"{got:.8f} ({got / want - 1:+.1e})")
        print(f"| {name} | " + " | ".join(cells) + " |")
    print("\nSpread between the four builds (max - min over builds, relative): " +
          ", ".join(f"gal {g}: {np.ptp([res[n]['scalars'][ib][g - 1][-1] for n, _ in VARIANTS]) / want:.1e}" for g, want in TUTORIAL.items()) + "\n")
    print("## All 196 example galaxies: row-relative difference to the pinned build (tests/parity.py metric)\n")
    print("Status characters / nsteps that differ from the pinned build are counted separately: a last-bit change can move "
          "a Brent iterate across its 1e-7 tolerance and with it `h`, `dt` and `nsteps`.\n")
    print("| build | quantity | median | 90 % | 99 % | max | rows > 1e-8 |")
    print("|---|---|---|---|---|---|---|")
    for name, _ in VARIANTS[1:]:
        r = res[name]
        for qi, q in enumerate(QUANT):
            valid = (base["profiles"][qi][..., 0] > -9e4) & (r["profiles"][qi][..., 0] > -9e4)
            e = row_err(r["profiles"][qi], base["profiles"][qi])[valid]
            if e.size == 0:
                continue
            print(f"| {name} | {q} | {np.median(e):.1e} | {np.quantile(e, 0.9):.1e} | {np.quantile(e, 0.99):.1e} | {e.max():.1e} | "
                  f"{int((e > 1e-8).sum())} of {e.size} |")
        ns = int((r["nsteps"] != base["nsteps"]).sum())
        st = int((np.asarray(r["status"]).view("S1") != np.asarray(base["status"]).view("S1")).sum())
        print(f"| {name} | nsteps / status | | | | | {ns} galaxies / {st} snapshots differ |")
    print("\n## Reading\n")
    print("* The three published tutorial values are reproduced by the pinned build to the residuals in the first row; the other "
          "builds -- same source, same algorithm, only the last bits of log/exp/pow/Bessel or the FMA contraction differ -- "
          "move the same three numbers by the amounts in the table.  A residual that is no larger than this build-to-build "
          "band cannot be told apart from the reference's own dependence on its (unpinned) libm / GSL / compiler flags.")
    print("* The amplification path is the one traced in DESIGN.md section 2: GSL's Brent solver (REL_TOL 1e-7, "
          "`root_finder.f90:171-184`) returns either the converged root or a bisection midpoint ~5e-8 away depending on "
          "the sign of a residual at rounding level; `xder` amplifies the kink in Omega_h into the shear, the pressure "
          "equation into `h`, `n`, and the dynamo into `Br`, `Bp`.")
    print("* This is why the 1e-8 bar of `north_star` is only meaningful between two implementations whose residual "
          "evaluations agree bit for bit, and why the CUDA path carries a textual twin of the pinned functions.")


if __name__ == "__main__":
    main()
