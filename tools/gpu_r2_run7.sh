#!/bin/bash
# round 2, GPU call 7: out-of-lined profile phase (construct_profiles, hydro_chain, write_record ... as functions): identity + time
mkdir -p gpurun_out
MAG_PROBE_LIBS=magnetizer_b200/libmag_ni.so timeout 600 python tools/r2_probe.py variants > gpurun_out/r2_probe7.log 2>&1; echo "rc=$?" >> gpurun_out/r2_probe7.log
cat gpurun_out/r2_probe7.log | cut -c1-220
