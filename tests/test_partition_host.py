"""CPU tests of the multi-GPU host logic (no GPU): the real csrc/mag_partition.cu linked with a CPU
fake of the single-GPU entry points (tests/partition_stub.cpp), driven through the same Python
mirror bench.py and the GPU tests use; and the world_size-2 (gloo) rendezvous of the
one-process-per-GPU launch, in which rank 0 drives the partition and the other rank only joins the
barriers."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")
STUB = os.path.join(BUILD, "libpartstub.so")
PROF, SCAL = ("Br", "h"), ("Bmax",)


def build_stub():
    os.makedirs(BUILD, exist_ok=True)
    srcs = [os.path.join(ROOT, "tests", "partition_stub.cpp"), os.path.join(ROOT, "magnetizer_b200", "csrc", "mag_partition.cu"),
            os.path.join(ROOT, "include", "magnetizer_b200.h")]
    if not os.path.exists(STUB) or any(os.path.getmtime(s) > os.path.getmtime(STUB) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-pthread", "-shared", "-o", STUB, srcs[0], "-x", "c++", srcs[1]])
    return STUB


@pytest.fixture()
def stub_solver():
    from magnetizer_b200 import solver as S
    old = (S._LIB, S.LIB_PATH)
    S._LIB, S.LIB_PATH = None, build_stub()
    yield S
    S._LIB, S.LIB_PATH = old


def expected(ids, inputs, nz, pq, sq):
    n = len(ids)
    i = np.arange(101) * 1e-3
    prof = np.stack([ids[:, None, None] * 1000.0 + q * 10.0 + inputs[0][:, :, None] + i[None, None, :] for q in pq])
    scal = np.stack([ids[:, None] + q * 0.5 + np.arange(nz)[None, :] * 0.01 for q in sq]) if len(sq) else np.zeros((0, n, nz))
    assert prof.shape == (len(pq), n, nz, 101)
    return prof, scal


def catalogue(n, nz=5, seed=1):
    rng = np.random.default_rng(seed)
    return np.arange(1, n + 1, dtype=np.int32), np.arange(nz, dtype=np.float64), rng.random((13, n, nz))


@pytest.mark.parametrize("ndev,n,chunk", [(1, 37, 4), (2, 37, 4), (3, 100, 7), (8, 1000, 16), (2, 5, 100), (4, 3, 1)])
def test_every_galaxy_solved_once_in_place(stub_solver, ndev, n, chunk):
    S = stub_solver
    from magnetizer_b200.abi import PROFILE_ID, SCALAR_ID
    ids, t, inp = catalogue(n)
    part = S.Partition(list(range(ndev)))
    try:
        r = part.solve(ids, t, inp, chunk=chunk, profiles=PROF, scalars=SCAL)
        st = part.stats()
    finally:
        part.close()
    nchunks = -(-n // chunk)
    assert r["chunks_done"].sum() == nchunks and st["chunk"] == chunk
    assert (r["chunks_stolen"] <= r["chunks_done"]).all()
    prof, scal = expected(ids, inp, t.size, [PROFILE_ID[q] for q in PROF], [SCALAR_ID[q] for q in SCAL])
    assert np.array_equal(r["profiles"], prof) and np.array_equal(r["scalars"], scal)
    assert np.array_equal(r["nsteps"], ids) and np.array_equal(r["point_stages"], ids.astype(np.int64) * 10)
    assert (r["status"] == b"M").all() and (r["error"] == 0).all()
    assert set(r["flags"].tolist()) <= set(range(ndev))          # the stub records the device that solved a galaxy
    assert st["launches"].sum() == 8 * nchunks


def test_static_share_and_stealing(stub_solver):
    """Chunk k belongs to device k % ndev; a device whose share is exhausted steals from the tails of the
    others.  With far more chunks than threads every device ends up with work, and nothing is stolen
    from a device that is listed alone."""
    S = stub_solver
    ids, t, inp = catalogue(400)
    part = S.Partition([0, 1, 2, 3])
    try:
        r = part.solve(ids, t, inp, chunk=5, profiles=PROF, scalars=SCAL)
        assert r["chunks_done"].sum() == 80 and (r["chunks_done"] > 0).all()
        one = S.Partition([0])
        try:
            r1 = one.solve(ids, t, inp, chunk=5, profiles=PROF, scalars=SCAL)
        finally:
            one.close()
        assert r1["chunks_done"].tolist() == [80] and r1["chunks_stolen"].tolist() == [0]
        assert np.array_equal(r["profiles"], r1["profiles"])
    finally:
        part.close()


def test_no_stealing_from_a_device_with_a_free_thread(stub_solver):
    """One chunk per device (the automatic chunk size of a strong-scaling run): the thread that loses the race for
    its own device's only chunk must not take the neighbour's before the neighbour's threads have started
    (round 2, first 8-GPU run: chunks done [2, 1, 0, 1, 0, 1, 1, 2])."""
    S = stub_solver
    for ndev in (2, 4, 8):
        ids, t, inp = catalogue(ndev * 40)
        part = S.Partition(list(range(ndev)))
        try:
            for _ in range(5):
                r = part.solve(ids, t, inp, chunk=40, profiles=PROF, scalars=SCAL)
                assert r["chunks_done"].tolist() == [1] * ndev and r["chunks_stolen"].tolist() == [0] * ndev
                assert sorted(set(r["flags"].tolist())) == list(range(ndev))      # every device solved its own chunk
                r = part.solve(ids, t, inp, chunk=20, profiles=PROF, scalars=SCAL)      # two chunks per device, two threads each
                assert r["chunks_done"].tolist() == [2] * ndev and r["chunks_stolen"].tolist() == [0] * ndev
        finally:
            part.close()


def test_automatic_chunk_size(stub_solver):
    S = stub_solver
    ids, t, inp = catalogue(5000, nz=2)
    part = S.Partition([0, 1])
    try:
        r = part.solve(ids, t, inp, chunk=0, profiles=("Br",), scalars=())
        assert part.stats()["chunk"] == 2500 and r["chunks_done"].sum() == 2     # one chunk per device
        big_ids, _, big_in = catalogue(60001, nz=1)
        part.solve(big_ids, t[:1], big_in, chunk=0, profiles=("Br",), scalars=())
        assert part.stats()["chunk"] == 15001                                    # 30 001 per device in two equal pieces
        part.solve(ids[:300], t, np.ascontiguousarray(inp[:, :300]), chunk=0, profiles=("Br",), scalars=())
        assert part.stats()["chunk"] == 256                                      # floor
    finally:
        part.close()


def test_resident_form_rows_stay_on_the_solving_device(stub_solver):
    S = stub_solver
    from magnetizer_b200.abi import PROFILE_ID
    n, nd = 90, 3
    ids, t, inp = catalogue(n)
    part = S.Partition(list(range(nd)))
    try:
        prof = [np.full((len(PROF), n, t.size, 101), -7.0) for _ in range(nd)]
        ps = [np.zeros(n, np.int64) for _ in range(nd)]
        owner = np.full(n, -1, np.int32)
        st = part.solve_device(n, t.size, [ids.ctypes.data] * nd, [t.ctypes.data] * nd, [inp.ctypes.data] * nd, PROF, (),
                               [dict(profiles=prof[d].ctypes.data, point_stages=ps[d].ctypes.data) for d in range(nd)],
                               chunk=6, owner=owner)
    finally:
        part.close()
    assert st["chunks_done"].sum() == 15 and (owner >= 0).all()
    want, _ = expected(ids, inp, t.size, [PROFILE_ID[q] for q in PROF], [])
    for g in range(n):
        for d in range(nd):
            if d == owner[g]:
                assert np.array_equal(prof[d][:, g], want[:, g]) and ps[d][g] == ids[g] * 10
            else:
                assert (prof[d][:, g] == -7.0).all() and ps[d][g] == 0


def test_error_of_one_chunk_fails_the_solve(stub_solver):
    S = stub_solver
    ids, t, inp = catalogue(64)
    ids = ids.copy()
    ids[40] = -1                      # the stub fails the chunk that starts with a negative id
    part = S.Partition([0, 1])
    try:
        with pytest.raises(S.MagnetizerError, match="poisoned"):
            part.solve(ids, t, inp, chunk=8, profiles=PROF, scalars=SCAL)
        ids[40] = 41                  # the partition object survives a failed solve
        r = part.solve(ids, t, inp, chunk=8, profiles=PROF, scalars=SCAL)
        assert r["chunks_done"].sum() == 8
    finally:
        part.close()


def test_create_fails_cleanly_on_a_missing_device(stub_solver):
    S = stub_solver
    with pytest.raises(S.MagnetizerError, match="no such device"):
        S.Partition([0, 100])


# ---------------------------------------------------------------------------------------------
def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from magnetizer_b200 import solver as S
    from magnetizer_b200.rendezvous import Rendezvous
    rv = Rendezvous(backend="gloo")
    assert rv.check_backend() == world
    if not rv.is_driver:
        rv.follow(3)
        q.put(("follower", rv.barriers))
        rv.close()
        return
    S._LIB, S.LIB_PATH = None, STUB
    ids, t, inp = catalogue(200)
    part = S.Partition(list(range(world)))          # rank 0 drives one (fake) device per rank
    rv.barrier()
    r = part.solve(ids, t, inp, chunk=10, profiles=PROF, scalars=SCAL)
    rv.barrier()
    part.close()
    q.put(("driver", rv.barriers + 1, r["chunks_done"].tolist(), bool((r["nsteps"] == ids).all())))
    rv.barrier()
    rv.close()


def test_two_rank_rendezvous_rank0_drives_the_partition():
    import torch.multiprocessing as mp
    build_stub()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180), q.get(timeout=180)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1] == ("follower", 3)
    assert got[0][0] == "driver" and got[0][1] == 3 and sum(got[0][2]) == 20 and got[0][3]
