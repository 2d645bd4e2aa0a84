#!/bin/bash
# round 2, GPU call 5: the two sweep cases again, ncu captures (launch list of a bench step, full sets of the three
# dominant kernels on bounded workloads), each only after its command has exited 0 without ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "no_Pdm or ignore_satellite or MAX_FAILS or driver_end_to_end" -s > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -4 gpurun_out/r2_pytest5.log
# launch list of one bench step at 25 000 galaxies (the partition's chunk size at N = 1)
timeout 600 python bench.py --steps 1 --warmup 3 --ngal 25000 --no-cpu-baseline > gpurun_out/r2_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --ngal 25000 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
# k_evolve on the bench MIX (synthetic galaxies truncated to 4 snapshots: the warps of an SM run different loop variants)
export MAG_HEAVY_MODE=0
timeout 300 python tools/ncu_workload_mix.py 2368 4 > gpurun_out/r2_plain_mix.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_evolve$ --launch-skip 1 -c 1 -o gpurun_out/prof_r2_evolve_mix \
    python tools/ncu_workload_mix.py 2368 4 > gpurun_out/r2_ncu_mix.log 2>&1
echo "k_evolve mix rc=$?"; cat gpurun_out/r2_plain_mix.log
# k_profiles on the same workload (second launch of the solve = pass 2 of the profile phase is -s 1; capture pass 1 and pass 2)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_profiles$ --launch-skip 2 -c 2 -o gpurun_out/prof_r2_profiles \
    python tools/ncu_workload_mix.py 2368 4 > gpurun_out/r2_ncu_profiles.log 2>&1
echo "k_profiles rc=$?"
# k_evolve_heavy on a uniform workload (444 copies of example galaxy 5, 3 snapshots)
export MAG_HEAVY_MODE=2
timeout 300 python tools/ncu_workload.py 444 5 3 > gpurun_out/r2_plain_heavy.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:^k_evolve_heavy$ --launch-skip 1 -c 1 -o gpurun_out/prof_r2_evolve_heavy \
    python tools/ncu_workload.py 444 5 3 > gpurun_out/r2_ncu_heavy.log 2>&1
echo "k_evolve_heavy rc=$?"; cat gpurun_out/r2_plain_heavy.log
ls -la gpurun_out/*.ncu-rep
