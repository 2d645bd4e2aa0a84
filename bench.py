#!/usr/bin/env python
"""bench.py -- galaxy histories / s of the per-galaxy dynamo path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ngal G] [--impl b200|reference]

Workload = BASELINE config 3: ONE synthetic catalogue of G = 1e5 Galform-shaped galaxy histories at
the default grid, solved on N GPUs (strong scaling: the catalogue is fixed, the GPUs share it).  A
"step" is one pass of the hot path over the whole catalogue through the multi-GPU partition of the
C ABI (mag_partition_*: static + work-stealing chunks, no collective on the solve path, results
gathered in place).  Under torchrun the ranks only rendezvous: rank 0 drives all N GPUs through the
partition -- the product's own multi-GPU form, one host process like the reference's master -- the
other ranks hold their barrier.

  value   catalogue resident in HBM (every GPU holds the inputs, rows stay on the GPU that solved
          them): mag_partition_solve_device, timed with CUDA events on every device, max over GPUs
  e2e     the reference-facing call with HOST buffers: mag_partition_solve, inputs from page-locked
          host memory, every result row stored into the caller's page-locked host arrays
  roofline  k_evolve + k_evolve_heavy of a single-context solve on GPU 0 (one stream, CUDA events)
  --impl reference  the CPU implementation of the same path (the oracle port of the reference's
          Fortran, which cannot be built here: no gfortran/MPI/HDF5/GSL) on all host cores, on a
          bounded sample of the same catalogue
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PROFILES = ("Br", "Bp", "alp_m", "h", "n")   # reduced output_quantities_list
SCALARS = ("Bmax",)
FLOP_PER_POINT_STAGE = 120.0                 # algorithmic, hoisted RHS
FP64_INSTR_PER_POINT_STAGE = 43.0            # FP64 instructions the loop issues per point and stage
METRIC = "galaxy histories/sec (FP64, default ngrid)"
CPU_SAMPLE = 192                             # galaxies of the CPU legs: every (G // 192)-th of the catalogue
WORKLOAD = "BASELINE config 3: synthetic Galform-shaped galaxy histories, default ngrid (p_nx_ref=101)"


def load_base():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    return d["t_gyr"], d["inputs"]


def base_config(G, nz):
    return {"workload": WORKLOAD, "galaxies_total": int(G), "snapshots": int(nz), "outputs": list(PROFILES) + list(SCALARS)}


def cpu_sample_ids(G):
    n = min(CPU_SAMPLE, G)
    return (np.arange(n, dtype=np.int64) * (G // n)).astype(np.int64)     # 0-based catalogue indices


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """CPU arm: the oracle port of the reference's Fortran path, all host threads, on a bounded sample
    (every (G // 192)-th galaxy) of the same catalogue the GPU arm solves."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_lib
    oracle_lib.build()
    from magnetizer_b200.synthetic import synthetic_subset
    t, base = load_base()
    G = args.ngal
    idx = cpu_sample_ids(G)
    inputs = synthetic_subset(base, idx)
    ids = (idx + 1).astype(np.int32)
    n = len(ids)
    p = oracle_lib.default_params()
    cores = os.cpu_count() or 1
    if args.warmup > 0:
        oracle_lib.solve_batch(p, ids[:cores], t, np.ascontiguousarray(inputs[:, :cores]), profiles=PROFILES, scalars=SCALARS, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_lib.solve_batch(p, ids, t, inputs, profiles=PROFILES, scalars=SCALARS, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    sample = (f"{n} galaxies = every {G // n}-th of the {G}-galaxy catalogue per step (the slice bench.py's GPU arm "
              f"checks against the oracle), {args.steps} step(s), {cores} threads with a shared work counter")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "galaxies/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": base_config(G, t.size),
        "cpu_baseline": {"value": v, "unit": "galaxies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "galaxies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def ncu_metrics():
    """ncu-derived numbers of the dominant kernel on the bench mix (committed summary, see profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_k_evolve_ncu.json")))
    except (OSError, ValueError):
        return {}


def run_b200(args):
    import torch
    from magnetizer_b200.rendezvous import Rendezvous
    from magnetizer_b200.solver import DynamoSolver, Partition, host_array, host_free
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories

    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rv = Rendezvous(backend="nccl", device=dev)      # NCCL carries one all-reduce at start-up, nothing on the solve path
    world = rv.world
    if world > 1:
        assert rv.check_backend(dev) == world

    def barrier():
        rv.barrier(sync=torch.cuda.synchronize)

    if not rv.is_driver:
        # the partition is driven from rank 0 (one host process, one thread pair per GPU); the other
        # ranks keep the rendezvous: around the timed regions and at the end
        rv.follow(5, sync=torch.cuda.synchronize)
        rv.close()
        return

    ndev = world
    if torch.cuda.device_count() < ndev:
        raise SystemExit(f"bench.py: {ndev} GPUs requested, {torch.cuda.device_count()} visible")
    devs = [torch.device("cuda", d) for d in range(ndev)]
    t, base = load_base()
    G, nz = args.ngal, t.size
    inputs = synthetic_histories(base, G)
    ids = galaxy_ids(G)
    part = Partition(list(range(ndev)))
    nr = part.params.p_nx_ref
    chunk = args.chunk

    # ---- HBM-resident arm: every GPU holds the catalogue's inputs and full-size output arrays
    res = []
    for d in devs:
        with torch.cuda.device(d):
            r = dict(inp=torch.from_numpy(inputs).to(d), t=torch.from_numpy(t).to(d), ids=torch.from_numpy(ids).to(d),
                     prof=torch.empty((len(PROFILES), G, nz, nr), dtype=torch.float64, device=d),
                     scal=torch.empty((len(SCALARS), G, nz), dtype=torch.float64, device=d),
                     status=torch.empty((G, nz), dtype=torch.uint8, device=d),
                     nsteps=torch.empty(G, dtype=torch.int32, device=d), error=torch.zeros(G, dtype=torch.int32, device=d),
                     ps=torch.zeros(G, dtype=torch.int64, device=d), flags=torch.empty(G, dtype=torch.int32, device=d),
                     flush=torch.empty(256 << 20, dtype=torch.uint8, device=d))   # > 126 MB L2
            res.append(r)
    d_outs = [dict(profiles=r["prof"].data_ptr(), scalars=r["scal"].data_ptr(), status=r["status"].data_ptr(),
                   nsteps=r["nsteps"].data_ptr(), error=r["error"].data_ptr(), point_stages=r["ps"].data_ptr(),
                   flags=r["flags"].data_ptr()) for r in res]
    owner = np.zeros(G, np.int32)

    def sync_all():
        for d in devs:
            torch.cuda.synchronize(d)

    def step_device():
        evs = []
        for d, r in zip(devs, res):
            with torch.cuda.device(d):
                r["flush"].fill_(1)
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
                evs.append([e0, None])
        sync_all()
        st = part.solve_device(G, nz, [r["ids"].data_ptr() for r in res], [r["t"].data_ptr() for r in res],
                               [r["inp"].data_ptr() for r in res], PROFILES, SCALARS, d_outs, chunk=chunk, owner=owner)
        for d, ev in zip(devs, evs):     # the call returns when every worker stream has drained
            with torch.cuda.device(d):
                ev[1] = torch.cuda.Event(enable_timing=True)
                ev[1].record()
        return evs, st

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(0)
    sampler.start()
    wall0 = time.perf_counter()
    runs = [step_device() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    # device time of a step = max over the GPUs of (end event - start event) on that GPU
    ms_dev = float(np.mean([max(a.elapsed_time(b) for a, b in evs) for evs, _ in runs]))
    st_last = runs[-1][1]
    pstats = part.stats()
    launches_step = int(pstats["launches"].sum())
    busy_ms = pstats["busy_ms"].tolist()
    W_total = 0
    n_err = 0
    for d in range(ndev):
        m = torch.from_numpy(owner == d).to(devs[d])
        W_total += int(res[d]["ps"][m].sum().item())
        n_err += int(res[d]["error"][m].sum().item())

    # ---- end-to-end arm: the reference-facing call, page-locked host buffers, results in place
    h_in = host_array(inputs.shape, np.float64)
    h_in[...] = inputs
    out = part.alloc_outputs(G, nz, PROFILES, SCALARS, pinned=True)
    e2e_steps = max(1, min(args.steps, 3))
    part.solve(ids, t, h_in, chunk=chunk, profiles=PROFILES, scalars=SCALARS, out=out)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        for d, r in zip(devs, res):
            with torch.cuda.device(d):
                r["flush"].fill_(1)
        sync_all()
        r_e2e = part.solve(ids, t, h_in, chunk=chunk, profiles=PROFILES, scalars=SCALARS, out=out)
    barrier()
    ms_e2e = (time.perf_counter() - w0) * 1e3 / e2e_steps
    e2e_stats = part.stats()
    h2d = inputs.nbytes + t.nbytes * int(r_e2e["chunks_done"].sum()) + ids.nbytes
    d2h = sum(out[k].nbytes for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"))
    assert int(r_e2e["point_stages"].sum()) == W_total, "e2e and HBM-resident arms disagree"

    # ---- roofline of the evolve kernels: one context, one stream, CUDA events (GPU 0, one chunk of the partition)
    # (the partition's contexts are released first: each holds a snapshot-record arena for its chunk, and a third one
    #  beside them would not fit a whole chunk and run it as two sub-launches)
    Gs = min(G, int(pstats["chunk"]))
    part.close()
    solver = DynamoSolver(device=0)
    with torch.cuda.device(devs[0]):
        r0 = res[0]
        sl_out = dict(profiles=r0["prof"].data_ptr(), scalars=r0["scal"].data_ptr(), status=r0["status"].data_ptr(),
                      nsteps=r0["nsteps"].data_ptr(), error=r0["error"].data_ptr(), point_stages=r0["ps"].data_ptr(),
                      flags=r0["flags"].data_ptr())
        # a contiguous slice [0, Gs) of the full arrays is itself a valid catalogue only for the per-galaxy
        # arrays; the [q][G][..] planes keep stride G, so solve the slice through the range form
        stream = torch.cuda.current_stream()
        for _ in range(2):
            r0["flush"].fill_(1)
            solver.run_device_range(G, 0, Gs, nz, r0["ids"].data_ptr(), r0["t"].data_ptr(), r0["inp"].data_ptr(), PROFILES,
                                    SCALARS, sl_out, stream=stream.cuda_stream)
            torch.cuda.synchronize()
        solver.device_status()
        tim = solver.last_timing()
        W_slice = int(r0["ps"][:Gs].sum().item())
        snaps_slice = int((r0["status"][:Gs] != ord("-")).sum().item())     # galaxy-snapshots the launch processed
        info = solver.info()
        assert info["records_chunk"] >= Gs, f"roofline leg ran as sub-launches of {info['records_chunk']} galaxies, not one launch of {Gs}"
        fp64_peak = solver.measure_fp64_peak(5)
    ms_evolve = tim["ms_kernels"] - tim["ms_profile_phase"]

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    flops = FLOP_PER_POINT_STAGE * W_slice / (ms_evolve * 1e-3)
    alg_bytes = h2d + d2h
    ncu = ncu_metrics()

    # ---- CPU baseline on the host cores: the oracle port on a bounded sample of the same catalogue,
    # and the GPU rows of those galaxies checked against it
    cpu = None
    if not args.no_cpu_baseline and ndev == 1:     # rank 0 at N = 1 only
        from oracle import oracle_lib
        oracle_lib.build()
        idx = cpu_sample_ids(G)
        cores = os.cpu_count() or 1
        sub_in = np.ascontiguousarray(inputs[:, idx])
        c0 = time.perf_counter()
        ora = oracle_lib.solve_batch(oracle_lib.default_params(), ids[idx], t, sub_in, profiles=PROFILES, scalars=SCALARS, nthreads=cores)
        cdt = time.perf_counter() - c0
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from parity import compare
        sub = {k: (np.asarray(v)[:, idx] if k in ("profiles", "scalars") else np.asarray(v)[idx]) if isinstance(v, np.ndarray) else v
               for k, v in r_e2e.items() if k not in ("chunks_done", "chunks_stolen")}
        rep = compare(sub, ora)
        cpu = {"value": len(idx) / cdt, "unit": "galaxies/s", "cores": cores, "kind": "port",
               "sample": f"{len(idx)} galaxies = every {G // len(idx)}-th of the catalogue, one pass, {cdt:.1f} s; the e2e arm's rows "
                         f"of those galaxies checked against it (worst row-relative error {max(rep.values()):.1e})"}

    cfg = base_config(G, nz)       # the same dict in both arms (--impl reference prints it too)
    partition = {"how": f"mag_partition_solve_device / mag_partition_solve: {ndev} GPU(s) x 2 host threads, chunks of {pstats['chunk']} "
                        "galaxies, static share + stealing, no collective; rank 0 drives every GPU",
                 "chunks_done": st_last["chunks_done"].tolist(), "chunks_stolen": st_last["chunks_stolen"].tolist(),
                 "gpu_busy_ms_per_step": [round(x, 1) for x in busy_ms],
                 "l2": "256 MiB flush write on every GPU between timed steps",
                 "point_stages_per_galaxy": W_total / G, "galaxies_with_error": n_err}
    line = {
        "metric": METRIC, "value": G / (ms_dev * 1e-3), "unit": "galaxies/s", "n_gpus": ndev,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "partition": partition,
        "e2e": {"value": G / (ms_e2e * 1e-3), "unit": "galaxies/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "steps": e2e_steps,
                "chunks_done": r_e2e["chunks_done"].tolist(), "chunks_stolen": r_e2e["chunks_stolen"].tolist(),
                "gpu_busy_ms_per_step": [round(x, 1) for x in e2e_stats["busy_ms"].tolist()]},
        "gpu_launches": int(launches_step * args.steps),
        "roofline": {"bound": "fp64", "achieved": flops / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                     "frac": flops / fp64_peak if fp64_peak else None,
                     # DRAM bytes of the launch: ncu's dram__bytes_read + dram__bytes_write per galaxy-snapshot (records read
                     # once, output rows written once, workspace write-backs) x the galaxy-snapshots of this launch
                     "traffic": (ncu["dram_bytes_per_galaxy_snapshot"] * snaps_slice) if "dram_bytes_per_galaxy_snapshot" in ncu else None,
                     "kernel": "k_evolve (+ k_evolve_heavy, concurrent)", "flop_per_point_stage": FLOP_PER_POINT_STAGE,
                     "point_stages_per_launch": W_slice, "galaxies_per_launch": Gs, "kernel_ms": ms_evolve,
                     "share_of_step": ms_evolve / tim["ms_kernels"], "profile_phase_ms": tim["ms_profile_phase"],
                     "heavy_galaxies": info["heavy_galaxies"], "fast_path_replays": info["fast_path_replays"],
                     # share of the FP64 pipe's issue slots that carry one of the 43 instructions of a LIVE point
                     "useful_slot_frac": (FP64_INSTR_PER_POINT_STAGE * W_slice / (ms_evolve * 1e-3)) / (fp64_peak / 2.0)
                     if fp64_peak else None,
                     "fp64_pipe_util": ncu.get("fp64_pipe_util"), "ncu_source": ncu.get("source"),
                     "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                     "hbm": {"achieved": alg_bytes / (ms_e2e * 1e-3) / 1e9, "peak": hbm_peak * ndev, "unit": "GB/s",
                             "frac": alg_bytes / (ms_e2e * 1e-3) / 1e9 / (hbm_peak * ndev),
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "wall_ms_per_step": wall * 1e3 / args.steps,
    }
    print(json.dumps(line))
    solver.close()
    part.close()
    host_free(h_in)
    for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"):
        host_free(out[k])
    barrier()
    rv.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ngal", type=int, default=int(os.environ.get("MAG_BENCH_NGAL", "100000")),
                    help="galaxies of the catalogue (total, shared by the GPUs): BASELINE config 3 = 1e5")
    ap.add_argument("--chunk", type=int, default=0, help="galaxies per partition chunk (0: four chunks per GPU)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
