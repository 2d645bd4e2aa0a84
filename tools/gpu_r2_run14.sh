#!/bin/bash
# round 2, one GPU, final state: the whole -m gpu suite, smoke(), the default bench line and the reference arm, then the C++ driver
# on 100 000 synthetic galaxies in four checkpoint cycles (the write of a cycle beside the solve of the next)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_n1_final.log 2>&1
echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_n1_final.log | cut -c1-3000
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_bench_ref_final.log 2>&1
echo "ref rc=$?"; tail -4 gpurun_out/r2_bench_ref_final.log | cut -c1-600
NGAL=100000
python tools/make_raw_catalogue.py $NGAL /tmp/c4.raw > gpurun_out/r2_driver_1gpu.log 2>&1
for mode in "" "--no-overlap"; do
  rm -f /tmp/c4_out.hdf5
  ( time timeout 900 magnetizer_b200/host/Magnetizer_b200 tools/config4_1gpu.in --raw-input /tmp/c4.raw --out /tmp/c4_out.hdf5 --restart --devices 0 $mode ) >> gpurun_out/r2_driver_1gpu.log 2>&1
  echo "driver $mode rc=$?" >> gpurun_out/r2_driver_1gpu.log
done
ls -la /tmp/c4_out.hdf5 >> gpurun_out/r2_driver_1gpu.log 2>&1
g++ -O1 -std=c++17 -pthread -o /tmp/h5probe tests/hdf5_writer_probe.cpp -lz && /tmp/h5probe rows /tmp/c4_out.hdf5 Output/Bmax $((NGAL - 3)) 3 > gpurun_out/r2_driver_rows.txt 2>&1
timeout 300 python - >> gpurun_out/r2_driver_1gpu.log 2>&1 <<PY
import numpy as np, sys
sys.path.insert(0, '.')
from magnetizer_b200.solver import DynamoSolver
from magnetizer_b200.synthetic import synthetic_subset
n = $NGAL
d = np.load('tests/golden/example_input.npz'); t, base = d['t_gyr'], d['inputs']
idx = np.arange(n - 3, n)
s = DynamoSolver(device=0)
r = s.run((idx + 1).astype(np.int32), t, synthetic_subset(base, idx), profiles=('Br',), scalars=('Bmax',))
rows = [float(x) for x in open('gpurun_out/r2_driver_rows.txt').read().split('\n')[1:] if x]
got = np.array(rows).reshape(3, t.size)
print('file rows of the last 3 galaxies vs a direct solve: max abs diff', float(np.abs(got - r['scalars'][0]).max()))
PY
grep -v "GPU 0\|status codes" gpurun_out/r2_driver_1gpu.log | tail -30
