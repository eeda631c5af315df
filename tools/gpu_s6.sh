#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dmma factorisation"; timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:4096 config3:4096 2>&1 | grep -v classes | tee gpurun_out/ab_solver_mma.log
echo "== generic factorisation"; MPC_RIC_GENERIC=1 timeout 600 python tools/ab_solver.py config2:4096 2>&1 | grep -v classes
timeout 900 python -m pytest tests -m gpu -x -q -k "forces_match or golden or extreme or status_codes or overflow" 2>&1 | tail -5 | tee gpurun_out/pytest_s6.log
