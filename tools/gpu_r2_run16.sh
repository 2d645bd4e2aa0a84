#!/bin/bash
# round 2, one GPU: memcheck of the fixed-physical-grid case, then the parity tests of the other new switches
mkdir -p gpurun_out
timeout 120 python tools/sanitize_case.py p_use_fixed_physical_grid=1 > gpurun_out/r2_fixedgrid_plain.log 2>&1; echo "plain rc=$?"; tail -3 gpurun_out/r2_fixedgrid_plain.log | cut -c1-300
( time timeout 500 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_case.py p_use_fixed_physical_grid=1 ) > gpurun_out/r2_fixedgrid_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -v "^$" gpurun_out/r2_fixedgrid_memcheck.log | head -60 | cut -c1-220
( time timeout 900 python -m pytest tests -m gpu -q -s -k "more_physics and not fixed_physical and not neumann" ) > gpurun_out/r2_pytest16.log 2>&1
echo "pytest rc=$?"; grep -n "worst\|passed\|failed\|Error" gpurun_out/r2_pytest16.log | tail -20 | cut -c1-250
