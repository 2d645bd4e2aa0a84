#!/bin/bash
# round 2, one GPU: the whole -m gpu suite, the default bench line, smoke(), the reference arm, and the launch list of the bench command
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_n1_final.log 2>&1
echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_n1_final.log | cut -c1-3000
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_bench_ref_final.log 2>&1
echo "ref rc=$?"; tail -4 gpurun_out/r2_bench_ref_final.log | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_launches_final.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_launches_final.log | cut -c1-300
