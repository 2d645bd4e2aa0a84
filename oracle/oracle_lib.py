"""ctypes loader of the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (magnetizer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from magnetizer_b200.abi import (MagOutputs, MagParams, PROFILE_ID, PROFILE_NAMES, SCALAR_ID,  # noqa: E402
                                 SCALAR_NAMES)

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("magnetizer_oracle.cpp", "oracle_numerics.hpp", "det_math.hpp", "bessel_coeffs.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "magnetizer_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_solve_batch.restype = C.c_int
        L.oracle_bessel.restype = C.c_double
        L.oracle_bessel.argtypes = [C.c_int32, C.c_double]
        L.oracle_constant.restype = C.c_double
        L.oracle_constant.argtypes = [C.c_int32]
        _LIB = L
    return _LIB


def default_params() -> MagParams:
    p = MagParams()
    lib().oracle_params_default(C.byref(p))
    return p


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def solve_batch(params, gal_ids, t_gyr, inputs, profiles=PROFILE_NAMES, scalars=SCALAR_NAMES, nthreads=0,
                want_hist=False):
    """Runs the oracle.  inputs: [13, ngal, nz].  Returns a dict of numpy arrays with the
    same layouts as the C ABI (profiles [nq, ngal, nz, nx_ref] ...)."""
    inputs = np.ascontiguousarray(inputs, np.float64)
    t_gyr = np.ascontiguousarray(t_gyr, np.float64)
    gal_ids = np.ascontiguousarray(gal_ids, np.int32)
    _, ngal, nz = inputs.shape
    assert gal_ids.shape == (ngal,) and t_gyr.shape == (nz,)
    pq = np.array([PROFILE_ID[q] for q in profiles], np.int32)
    sq = np.array([SCALAR_ID[q] for q in scalars], np.int32)
    nr = params.p_nx_ref
    res = dict(
        profiles=np.empty((len(pq), ngal, nz, nr), np.float64),
        scalars=np.empty((len(sq), ngal, nz), np.float64),
        status=np.empty((ngal, nz), "S1"),
        nsteps=np.empty(ngal, np.int32),
        error=np.empty(ngal, np.int32),
        point_stages=np.empty(ngal, np.int64),
        flags=np.empty(ngal, np.uint32),
    )
    out = MagOutputs(*[_ptr(res[k]) for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags")])
    hist_n = np.zeros((ngal, nz), np.int32) if want_hist else None
    hist_x = np.zeros((ngal, nz), np.int32) if want_hist else None
    rc = lib().oracle_solve_batch(C.byref(params), C.c_int32(ngal), _ptr(gal_ids), C.c_int32(nz), _ptr(t_gyr),
                                  _ptr(inputs), C.c_int32(len(pq)), _ptr(pq), C.c_int32(len(sq)), _ptr(sq),
                                  C.byref(out), C.c_int32(nthreads), _ptr(hist_n), _ptr(hist_x))
    if rc != 0:
        raise RuntimeError(f"oracle_solve_batch failed: {rc}")
    res["profile_names"] = tuple(profiles)
    res["scalar_names"] = tuple(scalars)
    if want_hist:
        res["nsteps_hist"], res["nx_hist"] = hist_n, hist_x
    return res
