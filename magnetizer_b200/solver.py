"""Host-side mirror of the reference operator boundary for the per-galaxy dynamo path.

Reference: ``runtime = dynamo_run(gal_id, test_run, error)`` (source/dynamo.f90:39) farmed
out by ``jobs_distribute`` (source/distributor.f90:180) and followed by
``write_output(gal_id, runtime)`` (source/output.f90:23).  Here: ``DynamoSolver.run`` solves
a batch of galaxy ids on one B200 through the C ABI (include/magnetizer_b200.h) and
``write_output_units`` applies the unit conversion table of write_output.

There is no CPU fallback: constructing a solver without the compiled CUDA library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import (ERROR_NAMES, MAG_NINPUT, MagOutputs, MagParams, PROFILE_ID, PROFILE_NAMES, SCALAR_ID,
                  SCALAR_NAMES)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAG_B200_LIB") or os.path.join(_HERE, "libmagnetizer_b200.so")
_LIB = None


class MagnetizerError(RuntimeError):
    pass


def load_library():
    """Loads the C-ABI shared library; raises if it was not built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise MagnetizerError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(make -C magnetizer_b200/csrc); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.mag_last_error.restype = C.c_char_p
        L.mag_measure_fp64_peak.restype = C.c_double
        L.mag_measure_fp64_peak.argtypes = [C.c_void_p, C.c_int]
        L.mag_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.mag_destroy.argtypes = [C.c_void_p]
        L.mag_destroy.restype = None
        L.mag_partition_last_error.restype = C.c_char_p
        L.mag_partition_destroy.argtypes = [C.c_void_p]
        L.mag_partition_destroy.restype = None
        L.mag_set_heavy_policy.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_int32]
        L.mag_host_alloc.restype = C.c_void_p
        L.mag_host_alloc.argtypes = [C.c_size_t]
        L.mag_host_free.argtypes = [C.c_void_p]
        L.mag_host_free.restype = None
        _LIB = L
    return _LIB


def default_params() -> MagParams:
    p = MagParams()
    load_library().mag_params_default(C.byref(p))
    return p


def _check(rc, what):
    if rc != 0:
        msg = load_library().mag_last_error().decode(errors="replace")
        raise MagnetizerError(f"{what} failed: {ERROR_NAMES.get(rc, rc)}: {msg}")


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class DynamoSolver:
    """One solver context on one GPU (replaces the per-rank module state of the reference)."""

    def __init__(self, params: MagParams | None = None, device: int = -1):
        self._lib = load_library()
        self.params = params if params is not None else default_params()
        self._ctx = C.c_void_p()
        _check(self._lib.mag_create(C.byref(self.params), device, C.byref(self._ctx)), "mag_create")

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.mag_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- diagnostics ---------------------------------------------------------
    def info(self):
        out = (C.c_int32 * 16)()
        _check(self._lib.mag_ctx_info(self._ctx, out), "mag_ctx_info")
        return dict(sm_count=out[0], profile_ctas_per_sm=out[1], profile_warps=out[2], nxcap=out[3], use_fast=out[4],
                    fast_path_replays=out[5], heavy_galaxies=out[6], evolve_teams_per_sm=out[7], evolve_teams=out[8],
                    heavy_teams_per_sm=out[9], heavy_teams=out[10], heavy_mode=out[11], records_chunk=out[12])

    def set_fast_path(self, on: bool):
        _check(self._lib.mag_set_fast_path(self._ctx, C.c_int32(1 if on else 0)), "mag_set_fast_path")

    def set_heavy_policy(self, mode: int = 1, alpha: float = 0.0, nx_heavy: int = 0):
        """Which galaxies run on four-warp teams (k_evolve_heavy): mode 0 none, 1 by cost, 2 all."""
        _check(self._lib.mag_set_heavy_policy(self._ctx, mode, float(alpha), nx_heavy), "mag_set_heavy_policy")

    def device_status(self):
        """Device-side status of the last run_device call (after the stream was synchronised)."""
        _check(self._lib.mag_last_device_status(self._ctx), "mag_last_device_status")

    def last_timing(self):
        ms = (C.c_double * 4)()
        n = self._lib.mag_last_timing(self._ctx, ms)
        return dict(kernel_launches=n, ms_profile_phase=ms[0], ms_kernels=ms[1], ms_h2d=ms[2], ms_d2h=ms[3])

    def measure_fp64_peak(self, repeats=5) -> float:
        return float(self._lib.mag_measure_fp64_peak(self._ctx, repeats))

    # -- the operator ----------------------------------------------------------
    def alloc_outputs(self, ngal, nz, profiles, scalars, pinned=False):
        nr = self.params.p_nx_ref

        def mk(shape, dt):
            if pinned:
                import torch
                tdt = {np.float64: torch.float64, np.int32: torch.int32, np.int64: torch.int64, np.uint32: torch.int32,
                       np.uint8: torch.uint8}[dt]
                return torch.empty(shape, dtype=tdt).pin_memory().numpy().view(dt)
            return np.empty(shape, dt)

        return dict(
            profiles=mk((len(profiles), ngal, nz, nr), np.float64),
            scalars=mk((len(scalars), ngal, nz), np.float64),
            status=mk((ngal, nz), np.uint8),
            nsteps=mk((ngal,), np.int32),
            error=mk((ngal,), np.int32),
            point_stages=mk((ngal,), np.int64),
            flags=mk((ngal,), np.uint32),
        )

    def run(self, gal_ids, t_gyr, inputs, profiles=PROFILE_NAMES, scalars=SCALAR_NAMES, out=None):
        """Solves the histories of `gal_ids` (1-based ids; only the RNG seed depends on them).

        inputs: float64 [13, ngal, nz] in the order of load_input_parameters
        (input_parameters.f90:158-170); t_gyr: [nz].  Returns a dict with
        profiles [nq, ngal, nz, nx_ref], scalars [ns, ngal, nz] (code units, as ts_data),
        status [ngal, nz] (S1), nsteps, error, point_stages, flags.
        """
        inputs = np.ascontiguousarray(inputs, np.float64)
        t_gyr = np.ascontiguousarray(t_gyr, np.float64)
        gal_ids = np.ascontiguousarray(gal_ids, np.int32)
        if inputs.ndim != 3 or inputs.shape[0] != MAG_NINPUT:
            raise ValueError("inputs must be [13, ngal, nz]")
        _, ngal, nz = inputs.shape
        if gal_ids.shape != (ngal,) or t_gyr.shape != (nz,):
            raise ValueError("gal_ids / t_gyr shape mismatch")
        pq = np.array([PROFILE_ID[q] for q in profiles], np.int32)
        sq = np.array([SCALAR_ID[q] for q in scalars], np.int32)
        res = out if out is not None else self.alloc_outputs(ngal, nz, profiles, scalars)
        o = MagOutputs(*[_ptr(res[k]) if res[k].size else None
                         for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags")])
        rc = self._lib.mag_solve_batch(self._ctx, C.c_int32(ngal), _ptr(gal_ids), C.c_int32(nz), _ptr(t_gyr),
                                       _ptr(inputs), C.c_int32(len(pq)), _ptr(pq), C.c_int32(len(sq)), _ptr(sq),
                                       C.byref(o))
        _check(rc, "mag_solve_batch")
        r = dict(res)
        r["status"] = res["status"].view("S1")
        r["profile_names"] = tuple(profiles)
        r["scalar_names"] = tuple(scalars)
        return r

    def run_device(self, ngal, nz, d_gal_ids, d_t_gyr, d_inputs, profiles, scalars, d_out: dict, stream=0):
        """Same with every buffer already resident in HBM (raw device pointers as ints);
        enqueues on `stream` (a cudaStream_t as int) and returns without synchronising."""
        self.run_device_range(ngal, 0, ngal, nz, d_gal_ids, d_t_gyr, d_inputs, profiles, scalars, d_out, stream)

    def run_device_range(self, ngal_total, g0, n, nz, d_gal_ids, d_t_gyr, d_inputs, profiles, scalars, d_out: dict, stream=0):
        """Galaxies [g0, g0+n) of full-catalogue device arrays, in place (mag_solve_range_device)."""
        pq = np.array([PROFILE_ID[q] for q in profiles], np.int32)
        sq = np.array([SCALAR_ID[q] for q in scalars], np.int32)
        o = MagOutputs(*[C.c_void_p(d_out[k]) if d_out.get(k) else None
                         for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags")])
        rc = self._lib.mag_solve_range_device(self._ctx, C.c_int32(ngal_total), C.c_int32(g0), C.c_int32(n), C.c_void_p(d_gal_ids),
                                              C.c_int32(nz), C.c_void_p(d_t_gyr), C.c_void_p(d_inputs), C.c_int32(len(pq)),
                                              _ptr(pq), C.c_int32(len(sq)), _ptr(sq), C.byref(o), C.c_void_p(stream))
        _check(rc, "mag_solve_range_device")


def host_array(shape, dtype):
    """numpy array in page-locked host memory that every GPU can store into (mag_host_alloc)."""
    lib = load_library()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    ptr = lib.mag_host_alloc(max(n, 1))
    if not ptr:
        raise MagnetizerError(f"mag_host_alloc({n}) failed")
    buf = (C.c_char * max(n, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    _HOST_ALLOCS[arr.ctypes.data] = ptr
    return arr


_HOST_ALLOCS: dict = {}


def host_free(arr):
    ptr = _HOST_ALLOCS.pop(arr.ctypes.data, None)
    if ptr:
        load_library().mag_host_free(ptr)


class Partition:
    """Static + work-stealing partition of a catalogue over the GPUs of one box
    (mag_partition_*; replaces jobs_distribute, distributor.f90:180): contexts and worker
    threads live as long as the object."""

    def __init__(self, devices, params: MagParams | None = None):
        self._lib = load_library()
        self.params = params if params is not None else default_params()
        self.devices = np.ascontiguousarray(devices, np.int32)
        self._h = C.c_void_p()
        rc = self._lib.mag_partition_create(C.byref(self.params), C.c_int32(self.devices.size), _ptr(self.devices),
                                            C.byref(self._h))
        self._check(rc, "mag_partition_create")

    def _check(self, rc, what):
        if rc != 0:
            raise MagnetizerError(f"{what} failed: {ERROR_NAMES.get(rc, rc)}: "
                                  f"{self._lib.mag_partition_last_error().decode(errors='replace')}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.mag_partition_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def alloc_outputs(self, ngal, nz, profiles, scalars, pinned=True):
        nr = self.params.p_nx_ref
        mk = host_array if pinned else (lambda shape, dt: np.empty(shape, dt))
        return dict(profiles=mk((len(profiles), ngal, nz, nr), np.float64), scalars=mk((len(scalars), ngal, nz), np.float64),
                    status=mk((ngal, nz), np.uint8), nsteps=mk((ngal,), np.int32), error=mk((ngal,), np.int32),
                    point_stages=mk((ngal,), np.int64), flags=mk((ngal,), np.uint32))

    def solve(self, gal_ids, t_gyr, inputs, chunk=0, profiles=PROFILE_NAMES, scalars=SCALAR_NAMES, out=None):
        """HOST arrays in, HOST arrays out (in place, galaxy order).  Returns the dict of
        DynamoSolver.run plus chunks_done / chunks_stolen per device."""
        inputs = np.ascontiguousarray(inputs, np.float64)
        t_gyr = np.ascontiguousarray(t_gyr, np.float64)
        gal_ids = np.ascontiguousarray(gal_ids, np.int32)
        _, ngal, nz = inputs.shape
        pq = np.array([PROFILE_ID[q] for q in profiles], np.int32)
        sq = np.array([SCALAR_ID[q] for q in scalars], np.int32)
        res = out if out is not None else self.alloc_outputs(ngal, nz, profiles, scalars, pinned=False)
        o = MagOutputs(*[_ptr(res[k]) if res[k].size else None
                         for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags")])
        done, stolen = np.zeros(self.devices.size, np.int32), np.zeros(self.devices.size, np.int32)
        rc = self._lib.mag_partition_solve(self._h, C.c_int32(ngal), _ptr(gal_ids), C.c_int32(nz), _ptr(t_gyr), _ptr(inputs),
                                           C.c_int32(len(pq)), _ptr(pq), C.c_int32(len(sq)), _ptr(sq), C.byref(o),
                                           C.c_int32(chunk), _ptr(done), _ptr(stolen))
        self._check(rc, "mag_partition_solve")
        r = dict(res)
        r["status"] = res["status"].view("S1")
        r["profile_names"], r["scalar_names"] = tuple(profiles), tuple(scalars)
        r["chunks_done"], r["chunks_stolen"] = done, stolen
        return r

    def stats(self):
        nd = self.devices.size
        chunk = C.c_int32()
        launches, busy = np.zeros(nd, np.int64), np.zeros(nd, np.float64)
        self._check(self._lib.mag_partition_stats(self._h, C.byref(chunk), _ptr(launches), _ptr(busy)), "mag_partition_stats")
        return dict(chunk=chunk.value, launches=launches, busy_ms=busy)

    def solve_device(self, ngal, nz, d_gal_ids, d_t_gyr, d_inputs, profiles, scalars, d_outs, chunk=0, owner=None):
        """Catalogue resident in HBM: per device raw pointers (ints) of its own copies of the inputs
        and of its own full-size output arrays (list of dicts as for DynamoSolver.run_device)."""
        nd = self.devices.size
        pq = np.array([PROFILE_ID[q] for q in profiles], np.int32)
        sq = np.array([SCALAR_ID[q] for q in scalars], np.int32)
        arr = lambda v: (C.c_void_p * nd)(*[C.c_void_p(x) for x in v])
        outs = (MagOutputs * nd)(*[MagOutputs(*[C.c_void_p(d[k]) if d.get(k) else None for k in
                                                ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags")])
                                   for d in d_outs])
        done, stolen = np.zeros(nd, np.int32), np.zeros(nd, np.int32)
        rc = self._lib.mag_partition_solve_device(self._h, C.c_int32(ngal), arr(d_gal_ids), C.c_int32(nz), arr(d_t_gyr),
                                                  arr(d_inputs), C.c_int32(len(pq)), _ptr(pq), C.c_int32(len(sq)), _ptr(sq),
                                                  outs, C.c_int32(chunk), _ptr(done), _ptr(stolen), _ptr(owner))
        self._check(rc, "mag_partition_solve_device")
        return dict(chunks_done=done, chunks_stolen=stolen)


def solve_partitioned(gal_ids, t_gyr, inputs, devices, chunk=0, params: MagParams | None = None,
                      profiles=PROFILE_NAMES, scalars=SCALAR_NAMES):
    """One-shot multi-GPU solve of one catalogue (mag_solve_partitioned = create + solve + destroy)."""
    part = Partition(devices, params)
    try:
        return part.solve(gal_ids, t_gyr, inputs, chunk=chunk, profiles=profiles, scalars=scalars)
    finally:
        part.close()


# ---------------------------------------------------------------------------
# write_output's unit table (reference source/output.f90:63-201; constants.f90:40-74).
def _units():
    km_kpc = 3.08567758e16
    s_gyr = 1.0e9 * 365.25 * 24 * 3600
    etat0_kmskpc = 1.0 / 3 * 0.1 * 10.0
    t0_kpcskm = 1.0 / etat0_kmskpc
    t0_s = t0_kpcskm * km_kpc
    h0_km = km_kpc
    h0_cm = h0_km * 1.0e5
    n0_cm3 = (1.0 / 1.0e6) ** 2 / h0_cm ** 2 * t0_s ** 2 / 1.67262158e-24
    vel = h0_km / t0_s  # h0_km/h0/t0_s*t0
    return dict(vel=vel, n=n0_cm3, t0_gyr=t0_kpcskm / s_gyr * km_kpc)


OUTPUT_UNITS = {  # quantity -> (factor name, unit string) as in write_output
    "Br": (1.0, "microgauss"), "Bp": (1.0, "microgauss"), "Bzmod": (1.0, "microgauss"), "Beq": (1.0, "microgauss"),
    "alp_m": ("vel", "km/s"), "alp_k": ("vel", "km/s"), "alp": ("vel", "km/s"), "v": ("vel", "km/s"),
    "Uz": ("vel", "km/s"), "h": (1.0e3, "pc"), "l": (1.0e3, "pc"), "Omega": ("vel", "km/s/kpc"),
    "Shear": ("vel", "km/s/kpc"), "etat": ("vel", "kpc km/s"), "n": ("n", "cm^-3"), "tau": (1.0, ""),
    "Omega_h": (1.0, "km/s/kpc"), "Omega_b": (1.0, "km/s/kpc"), "Omega_d": (1.0, "km/s/kpc"), "rkpc": (1.0, "kpc"),
    "rmax": (1.0, "kpc"), "Bmax": (1.0, "microgauss"), "Bmax_idx": (1.0, ""), "dt": (1.0, ""),
    "Bavg": (1.0, "microgauss"), "Beavg": (1.0, "microgauss"),
}


def write_output_units(name: str, values: np.ndarray) -> tuple[np.ndarray, str]:
    """Converts a ts_data quantity from code units to the units write_output stores."""
    fac, unit = OUTPUT_UNITS[name]
    if isinstance(fac, str):
        fac = _units()[fac]
    return values * fac, unit
