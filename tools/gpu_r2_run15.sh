#!/bin/bash
# round 2, one GPU: parity of the switches added last (Neumann outer BC, fixed physical grid, seeds, Dyn_quench = F)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x -s -k "more_physics or algebraic or tutorial or abi or driver_end_to_end" ) > gpurun_out/r2_pytest15.log 2>&1
echo "pytest rc=$?"; grep -v "^$" gpurun_out/r2_pytest15.log | tail -30 | cut -c1-250
