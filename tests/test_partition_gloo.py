"""world_size-2 (gloo, CPU) test of the multi-process host logic: static shards cover the
catalogue exactly once, the synthetic generator is shard-consistent, timing reduces as the
max over ranks and work as the sum, and the host-side gather restores catalogue order."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from magnetizer_b200.partition import gather_by_galaxy, reduce_step_stats, shard_bounds
    from magnetizer_b200.synthetic import synthetic_histories
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = np.load(os.path.join(ROOT, "tests", "golden", "example_input.npz"))["inputs"]
    total = 301
    g0, g1 = shard_bounds(total, world, rank)
    mine = synthetic_histories(base, g1 - g0, first=g0)
    stats = reduce_step_stats(100.0 + 50.0 * rank, g1 - g0, 1000 * (rank + 1), dist)
    counts = [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]
    rows = gather_by_galaxy(np.ascontiguousarray(mine[0]), counts, dist)  # r_disk rows, [n, nz]
    if rank == 0:
        full = synthetic_histories(base, total)
        q.put((stats, bool(np.array_equal(rows, full[0])), (g0, g1)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, gathered_ok, bounds = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stats == dict(ms_per_step=150.0, galaxies=301, point_stages=3000)
    assert gathered_ok and bounds == (0, 151)


def test_shard_bounds_cover_catalogue():
    sys.path.insert(0, ROOT)
    from magnetizer_b200.partition import shard_bounds
    for total, world in ((100000, 8), (7, 3), (5, 8)):
        b = [shard_bounds(total, world, r) for r in range(world)]
        assert b[0][0] == 0 and b[-1][1] == total
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
