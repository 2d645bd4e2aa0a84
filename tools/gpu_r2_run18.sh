#!/bin/bash
# round 2, one GPU, final state of the tree: the whole -m gpu suite and smoke()
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu_final2.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_final2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke2.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke2.log
