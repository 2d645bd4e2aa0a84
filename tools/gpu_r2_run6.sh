#!/bin/bash
# round 2, GPU call 6: ncu captures, exported to small CSV files on the box (the reports exceed gpurun's 64 MiB)
mkdir -p gpurun_out /tmp/ncu
timeout 600 python bench.py --steps 1 --warmup 3 --ngal 25000 --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --ngal 25000 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
# k_evolve on the bench MIX: synthetic galaxies, 6 snapshots each, at most 3000 steps per snapshot (no single history sets the time)
export MAG_HEAVY_MODE=0
timeout 300 python tools/ncu_workload_mix.py 4736 6 3000 > gpurun_out/r2_plain_mix.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_evolve$ --launch-skip 1 -c 1 -o /tmp/ncu/evolve_mix \
    python tools/ncu_workload_mix.py 4736 6 3000 > gpurun_out/r2_ncu_mix.log 2>&1
echo "k_evolve mix rc=$?"; cat gpurun_out/r2_plain_mix.log
python tools/ncu_export.py /tmp/ncu/evolve_mix.ncu-rep gpurun_out/r2_evolve_mix
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_profiles$ --launch-skip 2 -c 2 -o /tmp/ncu/profiles \
    python tools/ncu_workload_mix.py 4736 6 3000 > gpurun_out/r2_ncu_profiles.log 2>&1
echo "k_profiles rc=$?"
python tools/ncu_export.py /tmp/ncu/profiles.ncu-rep gpurun_out/r2_profiles
export MAG_HEAVY_MODE=2
timeout 300 python tools/ncu_workload.py 444 5 3 > gpurun_out/r2_plain_heavy.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:^k_evolve_heavy$ --launch-skip 1 -c 1 -o /tmp/ncu/evolve_heavy \
    python tools/ncu_workload.py 444 5 3 > gpurun_out/r2_ncu_heavy.log 2>&1
echo "k_evolve_heavy rc=$?"; cat gpurun_out/r2_plain_heavy.log
python tools/ncu_export.py /tmp/ncu/evolve_heavy.ncu-rep gpurun_out/r2_evolve_heavy
ls -la /tmp/ncu gpurun_out | tail -30
du -sh gpurun_out
