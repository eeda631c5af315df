#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:65536 2>&1 | grep -v "inverse \[" | tee gpurun_out/ab_solver_s11.log
for v in "MPC_RIC_BIG=0" "MPC_RIC_MCAP_BIG=16" "MPC_RIC_MCAP_BIG=32" "MPC_RIC_MCAP_BIG=16 MPC_RIC_128=1 MPC_RIC_MCAP=24"; do
  echo "== config3 $v" | tee -a gpurun_out/ab_solver_s11.log
  env $v timeout 600 python tools/ab_solver.py config3:4096 2>&1 | grep -v "inverse \[" | tee -a gpurun_out/ab_solver_s11.log
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_s11.log
python bench.py --config 1 --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/s11_c1.json
python tools/show_bench.py gpurun_out/s11_c1.json | grep -v parity
