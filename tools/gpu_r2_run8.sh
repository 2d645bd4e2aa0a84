#!/bin/bash
# round 2, 8-GPU call: BASELINE config 3 (1e5 galaxies, strong scaling) at N = 8, 4, 2 and config 4 (1e6 galaxies through the
# C++ driver: partition + HDF5 gather, write timed separately)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi8.txt 2>&1
nproc >> gpurun_out/r2_smi8.txt; free -g >> gpurun_out/r2_smi8.txt; df -h /tmp . >> gpurun_out/r2_smi8.txt
# the partition on eight physical GPUs against one context on GPU 0 (one-warp kernel only: bit-identical however the catalogue is cut)
MAG_HEAVY_MODE=0 timeout 600 python - > gpurun_out/r2_part8_check.log 2>&1 <<'PY'
import sys, time
import numpy as np
sys.path.insert(0, '.')
from magnetizer_b200.solver import DynamoSolver, Partition
from magnetizer_b200.synthetic import galaxy_ids, synthetic_histories
d = np.load('tests/golden/example_input.npz'); t, base = d['t_gyr'], d['inputs']
n = 4000
syn, ids = synthetic_histories(base, n, first=60000), galaxy_ids(n, first=60000)
prof, scal = ("Br", "Bp", "alp_m", "h", "n"), ("Bmax", "dt")
s = DynamoSolver(device=0); one = s.run(ids, t, syn, profiles=prof, scalars=scal); s.close()
import torch
nd = torch.cuda.device_count()
part = Partition(list(range(nd)))
out = part.alloc_outputs(n, t.size, prof, scal, pinned=True)
for chunk in (100, 250):
    t0 = time.perf_counter(); two = part.solve(ids, t, syn, chunk=chunk, profiles=prof, scalars=scal, out=out); dt = time.perf_counter() - t0
    same = all(np.array_equal(one[k], np.asarray(two[k])) for k in ("profiles", "scalars", "nsteps", "error", "point_stages", "flags")) and np.array_equal(one["status"], two["status"])
    print(f"{nd} GPUs, {n} galaxies, chunk {chunk}: {'bit-identical to one context' if same else 'DIFFERS'}; {dt*1e3:.0f} ms; chunks done {two['chunks_done'].tolist()} stolen {two['chunks_stolen'].tolist()}")
part.close()
PY
echo "partition check rc=$?"; cat gpurun_out/r2_part8_check.log | tail -4
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510 + N)) \
      bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/r2_bench_n$N.log 2>&1
  echo "bench N=$N rc=$?" | tee -a gpurun_out/r2_bench_n$N.log
  tail -2 gpurun_out/r2_bench_n$N.log | cut -c1-400
done
# config 4
FREE_GB=$(df --output=avail -BG /tmp | tail -1 | tr -dc 0-9)
NGAL=1000000; if [ "$FREE_GB" -lt 120 ]; then NGAL=250000; fi
echo "config 4: free /tmp ${FREE_GB} GB -> $NGAL galaxies" | tee gpurun_out/r2_config4.log
( time python tools/make_raw_catalogue.py $NGAL /tmp/c4.raw ) >> gpurun_out/r2_config4.log 2>&1
( time timeout 1500 magnetizer_b200/host/Magnetizer_b200 tools/config4.in --raw-input /tmp/c4.raw --out /tmp/config4_output.hdf5 --restart ) >> gpurun_out/r2_config4.log 2>&1
echo "driver rc=$?" >> gpurun_out/r2_config4.log
ls -la /tmp/config4_output.hdf5 >> gpurun_out/r2_config4.log 2>&1
g++ -O1 -std=c++17 -pthread -o /tmp/h5probe tests/hdf5_writer_probe.cpp -lz && /tmp/h5probe rows /tmp/config4_output.hdf5 Output/Bmax $((NGAL - 3)) 3 > gpurun_out/r2_config4_rows.txt 2>&1
timeout 300 python - >> gpurun_out/r2_config4.log 2>&1 <<PY
import numpy as np, sys
sys.path.insert(0, '.')
from magnetizer_b200.solver import DynamoSolver
from magnetizer_b200.synthetic import synthetic_subset
n = $NGAL
d = np.load('tests/golden/example_input.npz'); t, base = d['t_gyr'], d['inputs']
idx = np.arange(n - 3, n)
s = DynamoSolver(device=0)
r = s.run((idx + 1).astype(np.int32), t, synthetic_subset(base, idx), profiles=('Br',), scalars=('Bmax',))
rows = [float(x) for x in open('gpurun_out/r2_config4_rows.txt').read().split('\n')[1:] if x]
got = np.array(rows).reshape(3, t.size)
want = r['scalars'][0]
print('file rows of the last 3 galaxies vs a direct solve: max abs diff', float(np.abs(got - want).max()), 'shape', got.shape)
PY
tail -25 gpurun_out/r2_config4.log
