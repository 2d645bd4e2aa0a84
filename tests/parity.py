"""Parity checker: CUDA path vs CPU oracle on identical inputs (TEST INFRASTRUCTURE).

Bar (BASELINE.json north_star): status characters, nsteps, error flags and
Bmax_idx bit-exact; Br/Bp/alp_m and the derived ISM profiles within relative 1e-8 at
every output redshift.  "Relative" is taken per stored row (one quantity of one galaxy at
one snapshot, p_nx_ref points): max|a-b| <= tol * max|b|, so that zero crossings of Br/Bp
do not turn rounding noise into a spurious relative error.
"""
from __future__ import annotations

import numpy as np

RTOL = 1.0e-8
EXACT_SCALARS = ("Bmax_idx",)


def _row_err(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """max|a-b| / max|b| over the last axis (0 where both rows are identically equal)."""
    d = np.abs(a - b).max(axis=-1)
    s = np.abs(b).max(axis=-1)
    out = np.zeros_like(d)
    nz = d > 0
    out[nz] = d[nz] / np.where(s[nz] > 0, s[nz], 1e-300)
    return out


def compare(got: dict, want: dict, rtol: float = RTOL, check_cost: bool = True) -> dict:
    """Asserts parity of two result dicts (layouts of the C ABI); returns
    {quantity: worst row-relative error}."""
    assert got["profile_names"] == want["profile_names"] and got["scalar_names"] == want["scalar_names"]
    gs, ws = np.asarray(got["status"]).view("S1"), np.asarray(want["status"]).view("S1")
    if not np.array_equal(gs, ws):
        bad = np.argwhere(gs != ws)
        g = int(bad[0][0])
        raise AssertionError(f"status mismatch for galaxy index {g}: got {gs[g].tobytes()} want {ws[g].tobytes()} "
                             f"({len(bad)} entries differ)")
    for k in ("nsteps", "error"):
        assert np.array_equal(got[k], want[k]), (k, np.argwhere(got[k] != want[k])[:5].ravel())
    assert np.array_equal(got["flags"], want["flags"]), ("flags", np.argwhere(got["flags"] != want["flags"])[:5].ravel())
    if check_cost:
        assert np.array_equal(got["point_stages"], want["point_stages"]), "point_stages"
    rep = {}
    for qi, name in enumerate(want["profile_names"]):
        a, b = got["profiles"][qi], want["profiles"][qi]
        assert np.isfinite(a).all(), f"non-finite values in {name}"
        e = _row_err(a, b)
        rep[name] = float(e.max()) if e.size else 0.0
        if rep[name] > rtol:
            g, iz = np.unravel_index(np.argmax(e), e.shape)
            raise AssertionError(f"profile {name}: row-relative error {rep[name]:.3e} > {rtol} at galaxy index {g}, "
                                 f"snapshot {iz}")
    for qi, name in enumerate(want["scalar_names"]):
        a, b = got["scalars"][qi], want["scalars"][qi]
        if name in EXACT_SCALARS:
            assert np.array_equal(a, b), f"scalar {name} not exact"
            rep[name] = 0.0
            continue
        d = np.abs(a - b)
        s = np.abs(b)
        e = np.where(d > 0, d / np.where(s > 0, s, 1e-300), 0.0)
        rep[name] = float(e.max()) if e.size else 0.0
        if rep[name] > rtol:
            g, iz = np.unravel_index(np.argmax(e), e.shape)
            raise AssertionError(f"scalar {name}: relative error {rep[name]:.3e} > {rtol} at galaxy index {g}, "
                                 f"snapshot {iz}: {a[g, iz]!r} vs {b[g, iz]!r}")
    return rep
