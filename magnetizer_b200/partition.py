"""Host-side logic of the multi-process (one process per GPU) form of the solve.

Galaxies are independent: the catalogue is sharded over the ranks with no
data-path collective; torch.distributed is only used for the barrier around the timed
region and for reducing the timing (max over ranks) and the work counters (sum).  The
in-process multi-GPU form (one host thread per GPU, chunk stealing) is the C ABI's
mag_solve_partitioned (csrc/mag_partition.cu).  Replaces distributor.f90:253-384.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(ngal_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous static shard [g0, g1) of a catalogue of ngal_total galaxies."""
    base, rem = divmod(ngal_total, world)
    g0 = rank * base + min(rank, rem)
    return g0, g0 + base + (1 if rank < rem else 0)


def reduce_step_stats(ms_per_step: float, galaxies: int, point_stages: int, dist=None, device=None):
    """Whole-job statistics of one timed region: time = max over ranks, work = sum over ranks.
    Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(ms_per_step=ms_per_step, galaxies=galaxies, point_stages=point_stages)
    import torch
    t = torch.tensor([ms_per_step], dtype=torch.float64, device=device)
    w = torch.tensor([float(galaxies), float(point_stages)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return dict(ms_per_step=float(t[0]), galaxies=int(round(float(w[0]))), point_stages=int(round(float(w[1]))))


def gather_by_galaxy(local: np.ndarray, counts: list[int], dist=None) -> np.ndarray | None:
    """Host-side gather of per-rank result rows (first axis = galaxies of the rank's shard)
    into catalogue order on rank 0 (the writer).  Returns None on the other ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    src = torch.from_numpy(np.ascontiguousarray(local))
    if rank == 0:  # shards may differ in size by one galaxy: point-to-point, like the reference's MPI_Send/Recv
        parts = [src]
        for r in range(1, world):
            buf = torch.empty((counts[r],) + tuple(src.shape[1:]), dtype=src.dtype)
            dist.recv(buf, src=r)
            parts.append(buf)
        return torch.cat(parts, 0).numpy()
    dist.send(src, dst=0)
    return None
