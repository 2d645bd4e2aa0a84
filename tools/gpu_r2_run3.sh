#!/bin/bash
# round 2, GPU call 3: diagnose the two sweep failures, heavy-kernel variants, threshold
mkdir -p gpurun_out
timeout 600 python tools/r2_diag_sweep.py > gpurun_out/r2_diag3.log 2>&1; echo "diag rc=$?" >> gpurun_out/r2_diag3.log
tail -40 gpurun_out/r2_diag3.log
timeout 600 python tools/r2_probe.py cfg2 share mid tput > gpurun_out/r2_probe3_default.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe3_default.log
MAG_PROBE_ALPHAS=1.5,2.5 MAG_B200_LIB=$PWD/magnetizer_b200/libmag_t64.so timeout 600 python tools/r2_probe.py cfg2 share tput > gpurun_out/r2_probe3_t64.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe3_t64.log
MAG_PROBE_ALPHAS=1.5,2.5 MAG_B200_LIB=$PWD/magnetizer_b200/libmag_h4.so timeout 600 python tools/r2_probe.py cfg2 share tput > gpurun_out/r2_probe3_h4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe3_h4.log
grep -h "gal/s\|rc=" gpurun_out/r2_probe3_default.log gpurun_out/r2_probe3_t64.log gpurun_out/r2_probe3_h4.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "partition or heavy or multi_chunk or arena" > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -4 gpurun_out/r2_pytest3.log
