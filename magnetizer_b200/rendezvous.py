"""Rendezvous of the one-process-per-GPU launch (torchrun) around the multi-GPU partition.

The reference's MPI farm (distributor.f90:253-384) is replaced by ONE host process that drives
every GPU of the box through the C ABI's partition (mag_partition_*: one thread pair per GPU, no
collective on the solve path).  When the job is launched as N ranks, rank 0 is that process; the
other ranks only keep the rendezvous: they join the barriers around the timed regions and leave
with rank 0.  The default process group is whatever the launcher asks for ("nccl" on the GPUs,
"gloo" in the CPU tests); the barriers themselves always run on a gloo side group, so that a
waiting rank sleeps in the kernel instead of spinning a collective on its GPU while rank 0's
worker threads are using that GPU.
"""
from __future__ import annotations

import datetime
import os


class Rendezvous:
    def __init__(self, backend: str = "nccl", device=None, timeout_s: int = 7200):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self._dist = None
        self._side = None
        self.barriers = 0
        if self.world > 1:
            import torch.distributed as dist
            to = datetime.timedelta(seconds=timeout_s)
            kw = dict(device_id=device) if (backend == "nccl" and device is not None) else {}
            dist.init_process_group(backend, timeout=to, **kw)
            self._dist = dist
            self._side = dist.new_group(backend="gloo", timeout=to)

    @property
    def is_driver(self) -> bool:
        return self.rank == 0

    def check_backend(self, device=None) -> int:
        """One tiny all-reduce on the default group: every rank is up (returns the world size)."""
        if self._dist is None:
            return 1
        import torch
        one = torch.ones(1, device=device) if device is not None else torch.ones(1)
        self._dist.all_reduce(one)
        return int(one.item())

    def barrier(self, sync=None):
        if sync is not None:
            sync()
        if self._dist is not None:
            self._dist.barrier(group=self._side)
        self.barriers += 1

    def follow(self, n_barriers: int, sync=None):
        """What a rank other than the driver does: join the driver's barriers, nothing else."""
        for _ in range(n_barriers):
            self.barrier(sync)

    def close(self):
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None
