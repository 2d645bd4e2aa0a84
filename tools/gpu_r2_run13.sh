#!/bin/bash
# round 2, one GPU: fewer loop variants of the one-warp kernel (smaller instruction footprint) against the default build
mkdir -p gpurun_out
L=""; for v in G1 G2 G3 G4 K K2; do L="$L,magnetizer_b200/libmagv_$v.so"; done
MAG_PROBE_LIBS=magnetizer_b200/libmag_base.so$L MAG_PROBE_VARIANT_N=25000 MAG_PROBE_VARIANT_FIRST=40000 \
  timeout 900 python tools/r2_probe.py variants > gpurun_out/r2_probe_cands.log 2>&1
echo "probe rc=$?"; grep -v "^{" gpurun_out/r2_probe_cands.log | cut -c1-200
