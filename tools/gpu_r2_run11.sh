#!/bin/bash
# round 2, one GPU: the stage-loop rework of the one-warp kernel (no special-register read, no taken branch, own instantiation for
# grids that fill their last lane, sign-refresh step found ahead): parity subset, then A/B against the previous build
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "example_full_run or synthetic_histories or every_alignment or unstable or config5 or heavy_light or tutorial or generic_path" ) > gpurun_out/r2_pytest11.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest11.log
MAG_PROBE_LIBS=magnetizer_b200/libmag_base.so,magnetizer_b200/libmag_nodue.so MAG_PROBE_VARIANT_N=25000 MAG_PROBE_VARIANT_FIRST=40000 \
  timeout 600 python tools/r2_probe.py variants > gpurun_out/r2_probe_stage.log 2>&1
echo "probe rc=$?"; grep -v "^{" gpurun_out/r2_probe_stage.log | cut -c1-220
timeout 300 python tools/r2_probe.py share cfg2 > gpurun_out/r2_probe_stage2.log 2>&1
grep -v "^{" gpurun_out/r2_probe_stage2.log | cut -c1-220
