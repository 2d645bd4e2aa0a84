"""Seeded "Galform-shaped" synthetic galaxy histories (BASELINE configs 3-5).

The reference's inputs come from Galform merger trees (python/prepare_input.py:93-150 of
the reference defines the contract: 13 time series per galaxy, negative sentinels outside a
galaxy's life).  No Galform output is available offline, so synthetic galaxies are
bootstrapped from a set of base histories (the 196 x 40 example input): galaxy g copies
base history perm[g mod nbase] and multiplies each column by one per-galaxy log-normal
factor, which preserves temporal correlation, validity windows, sentinel signs, elliptical
phases and disk-growth statistics.
"""
from __future__ import annotations

import numpy as np

# column order of load_input_parameters (input_parameters.f90:158-170)
_SIGMA = np.array([0.10, 0.10, 0.10, 0.10, 0.10, 0.10, 0.0, 0.30, 0.30, 0.30, 0.30, 0.30, 0.0])
SEED = 20240607


def synthetic_histories(base_inputs: np.ndarray, ngal: int, seed: int = SEED, first: int = 0) -> np.ndarray:
    """Returns inputs [13, ngal, nz] for synthetic galaxies first .. first+ngal-1.

    Galaxy g's history depends on (seed, g) only, so any shard of a larger catalogue can be
    generated independently (used for the multi-GPU partition).
    """
    ncol, nbase, nz = base_inputs.shape
    assert ncol == 13
    perm = np.random.Generator(np.random.PCG64(seed)).permutation(nbase)
    g = np.arange(first, first + ngal, dtype=np.int64)
    src = perm[g % nbase]
    out = np.ascontiguousarray(base_inputs[:, src, :])
    # one factor per (galaxy, column): counter-based so that shards agree
    fac = np.empty((ncol, ngal))
    blk = 4096
    for b0 in range(first // blk * blk, first + ngal, blk):
        rng = np.random.Generator(np.random.PCG64([seed, b0 // blk]))
        z = rng.standard_normal((blk, ncol))
        lo, hi = max(b0, first), min(b0 + blk, first + ngal)
        fac[:, lo - first:hi - first] = np.exp(_SIGMA[None, :] * z[lo - b0:hi - b0, :]).T
    # the first nbase galaxies of the catalogue are the base histories themselves
    fac[:, (g < nbase)] = 1.0
    out *= fac[:, :, None]
    return out


def synthetic_subset(base_inputs: np.ndarray, indices, seed: int = SEED) -> np.ndarray:
    """inputs [13, len(indices), nz] of the catalogue galaxies with the given 0-based indices
    (the same histories synthetic_histories(base, G)[:, indices] holds for any G > max index)."""
    out = [synthetic_histories(base_inputs, 1, seed=seed, first=int(i)) for i in indices]
    return np.ascontiguousarray(np.concatenate(out, axis=1))


def galaxy_ids(ngal: int, first: int = 0) -> np.ndarray:
    """1-based ids (the RNG seed of a galaxy depends on its id, random.f90:43)."""
    return np.arange(first + 1, first + ngal + 1, dtype=np.int32)
