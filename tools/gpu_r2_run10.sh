#!/bin/bash
# round 2, one GPU: ncu --set full of k_evolve on the bench's own workload (one 25 000-galaxy chunk, full histories), every SASS
# line exported; then the heavy/light split at finer alpha on a 12 500-galaxy share
mkdir -p gpurun_out
export MAG_HEAVY_MODE=0
timeout 300 python tools/ncu_workload_mix.py 25000 40 0 40000 > gpurun_out/r2_plain_chunk.log 2>&1
echo "plain rc=$?"; cat gpurun_out/r2_plain_chunk.log | tail -2
( time timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_evolve$ --launch-skip 1 -c 1 -o /tmp/prof_r2_evolve_chunk \
    python tools/ncu_workload_mix.py 25000 40 0 40000 ) > gpurun_out/r2_ncu_chunk.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/r2_ncu_chunk.log
python tools/ncu_export.py /tmp/prof_r2_evolve_chunk.ncu-rep gpurun_out/r2_evolve_chunk 0
ls -la gpurun_out/r2_evolve_chunk*
unset MAG_HEAVY_MODE
MAG_PROBE_ALPHAS=1.25,1.5,1.75 timeout 600 python tools/r2_probe.py share > gpurun_out/r2_probe_alpha.log 2>&1
echo "probe rc=$?"; tail -8 gpurun_out/r2_probe_alpha.log | cut -c1-200
