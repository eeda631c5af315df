#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:65536 config3:4096 2>&1 | grep -v "classes\|riccati \[" | tee gpurun_out/ab_solver_s8.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_s8.log
B="python bench.py --steps 3 --warmup 3 --no-cpu --config 2"
ncu --clock-control none --set full --import-source on --kernel-name-base demangled -k regex:riccati -s 9 -c 3 -f -o gpurun_out/r2_ric_c2 $B > gpurun_out/r2_ric_c2.log 2>&1
LIB=quadruped_ctrl_b200/libquadruped_mpc_b200.so
python tools/ncu_summary.py gpurun_out/r2_ric_c2.ncu-rep smsp__average_warps_issue_stalled sm__pipe_tensor sm__inst_executed_pipe > gpurun_out/r2_ric_c2_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_ric_c2.ncu-rep $LIB riccati 700 > gpurun_out/r2_ric_c2_lines.txt 2>&1
rm -f gpurun_out/r2_ric_c2.ncu-rep
python bench.py --config 2 --steps 100 --warmup 5 --no-cpu 2>gpurun_out/b2.err | tail -1 > gpurun_out/s8_c2.json
python tools/show_bench.py gpurun_out/s8_c2.json | grep -v parity
