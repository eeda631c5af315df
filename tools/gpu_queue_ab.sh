#!/bin/bash
# A/B of the dynamic problem queue against the static stride (MPC_RIC_STATIC=1), one batch at a time, both solvers.
set -u
mkdir -p gpurun_out
echo "== dynamic queue"; timeout 600 python tools/ab_solver.py config2:4096 config5:65536 four_stance:4096 config3:4096 2>&1 | grep -v "inverse \[\|riccati \["
echo "== static stride"; MPC_RIC_STATIC=1 timeout 600 python tools/ab_solver.py config2:4096 config5:65536 four_stance:4096 config3:4096 2>&1 | grep -v "inverse \[\|riccati \["
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --config 3 --steps 50 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/q_c3.json; python tools/show_bench.py gpurun_out/q_c3.json | grep -v parity
