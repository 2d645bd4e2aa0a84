#!/usr/bin/env python
"""bench.py -- galaxy histories / s of the per-galaxy dynamo path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--ngal G] [--impl b200|reference]

A "step" is one pass of the hot path (mag_solve_batch) over one batch of G synthetic
Galform-shaped galaxy histories PER GPU (weak scaling; galaxies are independent, there is
no collective on the solve path).  `value` times the pass with inputs resident in HBM
(mag_solve_batch_device, CUDA events on the launching stream); `e2e` times the
reference-facing call with pinned HOST buffers (H2D + kernels + D2H inside the region).
`--impl reference` times the CPU implementation of the same path (the oracle port of the
reference's Fortran, which cannot be built here: no gfortran/MPI/HDF5/GSL) on all host
cores on a bounded sample.
"""
from __future__ import annotations

import argparse
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
METRIC = "galaxy histories/sec (FP64, default ngrid)"
CPU_SAMPLE = 196


def load_base():
    d = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))
    return d["t_gyr"], d["inputs"]


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
    """CPU arm: the oracle port of the reference's Fortran path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle_lib
    oracle_lib.build()
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
    t, base = load_base()
    n = CPU_SAMPLE
    inputs = synthetic_histories(base, n)
    ids = galaxy_ids(n)
    p = oracle_lib.default_params()
    cores = os.cpu_count() or 1
    for _ in range(args.warmup if args.warmup < 1 else 1):
        oracle_lib.solve_batch(p, ids[:cores], t, inputs[:, :cores], profiles=PROFILES, scalars=SCALARS, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_lib.solve_batch(p, ids, t, inputs, profiles=PROFILES, scalars=SCALARS, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    sample = f"first {n} galaxies of the synthetic catalogue (= the 196 example histories), {args.steps} pass(es)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "galaxies/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic Galform-shaped galaxy histories, default ngrid (p_nx_ref=101)",
                   "galaxies_per_step": n},
        "cpu_baseline": {"value": v, "unit": "galaxies/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "galaxies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist
    from magnetizer_b200.abi import PROFILE_ID  # noqa: F401
    from magnetizer_b200.solver import DynamoSolver
    from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t, base = load_base()
    G = args.ngal
    nz = t.size
    inputs = synthetic_histories(base, G, first=rank * G)     # this rank's shard of the catalogue
    ids = galaxy_ids(G, first=rank * G)
    solver = DynamoSolver(device=local)
    nr = solver.params.p_nx_ref

    # ---- HBM-resident arm
    d_in = torch.from_numpy(inputs).to(dev)
    d_t = torch.from_numpy(t).to(dev)
    d_ids = torch.from_numpy(ids).to(dev)
    d_prof = torch.empty((len(PROFILES), G, nz, nr), dtype=torch.float64, device=dev)
    d_scal = torch.empty((len(SCALARS), G, nz), dtype=torch.float64, device=dev)
    d_status = torch.empty((G, nz), dtype=torch.uint8, device=dev)
    d_nsteps = torch.empty(G, dtype=torch.int32, device=dev)
    d_error = torch.empty(G, dtype=torch.int32, device=dev)
    d_ps = torch.empty(G, dtype=torch.int64, device=dev)
    d_flags = torch.empty(G, dtype=torch.int32, device=dev)
    d_out = dict(profiles=d_prof.data_ptr(), scalars=d_scal.data_ptr(), status=d_status.data_ptr(),
                 nsteps=d_nsteps.data_ptr(), error=d_error.data_ptr(), point_stages=d_ps.data_ptr(),
                 flags=d_flags.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        solver.run_device(G, nz, d_ids.data_ptr(), d_t.data_ptr(), d_in.data_ptr(), PROFILES, SCALARS, d_out,
                          stream=stream.cuda_stream)
        e1.record(stream)
        return e0, e1

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    wall0 = time.perf_counter()
    evs = [step_device() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if sampler else None
    ms_dev = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    tim_dev = solver.last_timing()
    launches = tim_dev["kernel_launches"]
    ms_evolve = tim_dev["ms_kernels"] - tim_dev["ms_profile_phase"]   # k_evolve of the last timed step
    W_rank = int(d_ps.sum().item())
    n_err = int(d_error.sum().item())

    # ---- end-to-end arm: reference-facing call, pinned host buffers
    h_in = torch.from_numpy(inputs).pin_memory().numpy()
    out = solver.alloc_outputs(G, nz, PROFILES, SCALARS, pinned=True)
    e2e_steps = max(1, min(args.steps, 3))
    solver.run(ids, t, h_in, profiles=PROFILES, scalars=SCALARS, out=out)
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        flush.fill_(1)
        r = solver.run(ids, t, h_in, profiles=PROFILES, scalars=SCALARS, out=out)
    barrier()
    ms_e2e = (time.perf_counter() - w0) * 1e3 / e2e_steps
    tim = solver.last_timing()
    h2d = inputs.nbytes + t.nbytes + ids.nbytes
    d2h = sum(out[k].nbytes for k in ("profiles", "scalars", "status", "nsteps", "error", "point_stages", "flags"))
    assert np.array_equal(r["point_stages"], d_ps.cpu().numpy()), "e2e and HBM-resident arms disagree"

    red = torch.tensor([ms_dev, ms_e2e, wall * 1e3 / args.steps], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(W_rank), float(G)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_dev_max, ms_e2e_max, ms_wall_max = red.tolist()
    W_all, G_all = tot.tolist()

    if rank == 0:
        fp64_peak = solver.measure_fp64_peak(5)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        flops = FLOP_PER_POINT_STAGE * W_rank / (ms_evolve * 1e-3)
        alg_bytes = h2d + d2h
        # CPU baseline on rank 0: the oracle port on a bounded sample of the same workload
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle_lib
            oracle_lib.build()
            n = min(CPU_SAMPLE, G)
            cores = os.cpu_count() or 1
            c0 = time.perf_counter()
            ora = oracle_lib.solve_batch(oracle_lib.default_params(), ids[:n], t, np.ascontiguousarray(inputs[:, :n]),
                                         profiles=PROFILES, scalars=SCALARS, nthreads=cores)
            cdt = time.perf_counter() - c0
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            from parity import compare
            sub = {k: (v[:, :n] if k in ("profiles", "scalars") else v[:n]) if isinstance(v, np.ndarray) else v
                   for k, v in r.items()}
            rep = compare(sub, ora)
            cpu = {"value": n / cdt, "unit": "galaxies/s", "cores": cores, "kind": "port",
                   "sample": f"first {n} galaxies of rank 0's batch, one pass, {cdt:.1f} s; GPU result checked against it "
                             f"(worst row-relative error {max(rep.values()):.1e})"}
        line = {
            "metric": METRIC, "value": G_all / (ms_dev_max * 1e-3), "unit": "galaxies/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic Galform-shaped galaxy histories, default ngrid (p_nx_ref=101)",
                       "galaxies_per_gpu": G, "galaxies_total": int(G_all), "snapshots": int(nz),
                       "outputs": list(PROFILES) + list(SCALARS), "partition": f"static x{world}, no collective",
                       "l2": "256 MiB flush write between timed steps",
                       "point_stages_per_galaxy": W_all / G_all, "galaxies_with_error": n_err},
            "e2e": {"value": G_all / (ms_e2e_max * 1e-3), "unit": "galaxies/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e_max,
                    "ms_h2d": tim["ms_h2d"], "ms_kernels": tim["ms_kernels"], "ms_profile_phase": tim["ms_profile_phase"],
                    "ms_d2h": tim["ms_d2h"]},
            "gpu_launches": int(launches * args.steps),
            "roofline": {"bound": "fp64", "achieved": flops / 1e12, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                         "frac": flops / fp64_peak if fp64_peak else None, "traffic": None,
                         "kernel": "k_evolve", "flop_per_point_stage": FLOP_PER_POINT_STAGE,
                         "point_stages_per_launch": W_rank, "kernel_ms": ms_evolve,
                         "share_of_step": ms_evolve / ms_dev, "profile_phase_ms": tim_dev["ms_profile_phase"],
                         "fast_path_replays": solver.info()["fast_path_replays"],
                         "peak_source": "DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "hbm": {"achieved": alg_bytes / (ms_dev * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes / (ms_dev * 1e-3) / 1e9 / hbm_peak,
                                 "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "wall_ms_per_step": ms_wall_max,
        }
        print(json.dumps(line))
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--ngal", type=int, default=int(os.environ.get("MAG_BENCH_NGAL", "40000")),
                    help="galaxies per GPU per step (weak scaling).  The workload is BASELINE config 3's synthetic "
                         "catalogue; the default batch is large enough that the longest history (800 000 RK steps, "
                         "~1.8 s on its warp) does not set the step time")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
