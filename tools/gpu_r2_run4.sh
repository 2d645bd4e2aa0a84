#!/bin/bash
# round 2, GPU call 4: full GPU suite, profile-phase register variants, partition threads/chunks, default bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -6 gpurun_out/r2_pytest4.log
: > gpurun_out/r2_probe4.log
for lib in libmagnetizer_b200.so libmag_p96.so libmag_p168.so libmag_p255.so; do
  echo "== $lib" >> gpurun_out/r2_probe4.log
  MAG_B200_LIB=$PWD/magnetizer_b200/$lib timeout 300 python tools/r2_probe.py prof >> gpurun_out/r2_probe4.log 2>&1
done
timeout 600 python tools/r2_probe.py part >> gpurun_out/r2_probe4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe4.log
grep -h "==\|gal/s\|rc=" gpurun_out/r2_probe4.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench4.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2_bench4.log
tail -3 gpurun_out/r2_bench4.log
