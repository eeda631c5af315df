#!/bin/bash
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_solver_s9.log
for m in default 6 10 16; do
  echo "== MPC_RIC_MCAP=$m" | tee -a gpurun_out/ab_solver_s9.log
  if [ $m = default ]; then unset MPC_RIC_MCAP; else export MPC_RIC_MCAP=$m; fi
  timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:65536 config3:4096 2>&1 | grep -v "classes" | tee -a gpurun_out/ab_solver_s9.log
done
unset MPC_RIC_MCAP
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_s9.log
