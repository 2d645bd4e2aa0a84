#!/bin/bash
# round 2, GPU call 1: full GPU test suite, kernel split probes, build-time variants, first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -s > gpurun_out/r2_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
MAG_PROBE_LIBS=magnetizer_b200/libmag_xch.so,magnetizer_b200/libmag_b2.so,magnetizer_b200/libmag_noinl.so \
  timeout 900 python tools/r2_probe.py cfg2 share tput variants > gpurun_out/r2_probe1.log 2>&1
echo "probe rc=$?" >> gpurun_out/r2_probe1.log
tail -40 gpurun_out/r2_probe1.log
timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/r2_bench1.log 2>&1
echo "bench rc=$?" >> gpurun_out/r2_bench1.log
tail -3 gpurun_out/r2_bench1.log
