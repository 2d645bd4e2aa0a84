#!/bin/bash
# round 2, 2-GPU call: the partition on two physical GPUs against one context, bench at N = 2
mkdir -p gpurun_out
MAG_HEAVY_MODE=0 timeout 600 python - > gpurun_out/r2_part2_check.log 2>&1 <<'PY'
import sys, time
import numpy as np
sys.path.insert(0, '.')
from magnetizer_b200.solver import DynamoSolver, Partition
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
import torch
d = np.load('tests/golden/example_input.npz'); t, base = d['t_gyr'], d['inputs']
n = 4000
syn, ids = synthetic_histories(base, n, first=60000), galaxy_ids(n, first=60000)
prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
s = DynamoSolver(device=0); one = s.run(ids, t, syn, profiles=prof, scalars=scal); s.close()
nd = torch.cuda.device_count()
part = Partition(list(range(nd)))
out = part.alloc_outputs(n, t.size, prof, scal, pinned=True)
for chunk in (100, 500):
    t0 = time.perf_counter(); two = part.solve(ids, t, syn, chunk=chunk, profiles=prof, scalars=scal, out=out); dt = time.perf_counter() - t0
    same = all(np.array_equal(one[k], np.asarray(two[k])) for k in ("profiles", "scalars", "nsteps", "error", "point_stages", "flags")) and np.array_equal(one["status"], two["status"])
    print(f"{nd} GPUs, {n} galaxies, chunk {chunk}: {'bit-identical to one context' if same else 'DIFFERS'}; {dt*1e3:.0f} ms; chunks done {two['chunks_done'].tolist()} stolen {two['chunks_stolen'].tolist()}")
part.close()
PY
echo "partition check rc=$?"; cat gpurun_out/r2_part2_check.log | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2.log 2>&1
echo "bench N=2 rc=$?" | tee -a gpurun_out/r2_bench_n2.log
tail -3 gpurun_out/r2_bench_n2.log | cut -c1-600
