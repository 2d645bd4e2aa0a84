#!/bin/bash
# round 2, GPU call 2: full GPU test suite, policy probes, first full-size bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
timeout 900 python tools/r2_probe.py cfg2 share tput > gpurun_out/r2_probe2.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2_probe2.log
tail -45 gpurun_out/r2_probe2.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench2.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2_bench2.log
tail -3 gpurun_out/r2_bench2.log
