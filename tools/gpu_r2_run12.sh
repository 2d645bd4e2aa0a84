#!/bin/bash
# round 2, one GPU: one library per stage-loop change of the one-warp kernel (tools/README.md), each against the default build
mkdir -p gpurun_out
L=""; for v in none A B C D E F; do L="$L,magnetizer_b200/libmagv_$v.so"; done
MAG_PROBE_LIBS=magnetizer_b200/libmag_base.so$L MAG_PROBE_VARIANT_N=25000 MAG_PROBE_VARIANT_FIRST=40000 \
  timeout 900 python tools/r2_probe.py variants > gpurun_out/r2_probe_toggles.log 2>&1
echo "probe rc=$?"; grep -v "^{" gpurun_out/r2_probe_toggles.log | cut -c1-200
