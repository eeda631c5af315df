#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:4096 config5:65536 config3:4096 2>&1 | grep -v classes | tee gpurun_out/ab_solver_fast.log
timeout 900 python -m pytest tests -m gpu -x -q -k "forces_match or golden or extreme or status_codes or overflow" 2>&1 | tail -5 | tee gpurun_out/pytest_s7.log
B="python bench.py --steps 3 --warmup 3 --no-cpu --config 2"
ncu --clock-control none --set full --import-source on --kernel-name-base demangled -k regex:riccati -s 9 -c 3 -f -o gpurun_out/r2_ric_c2 $B > gpurun_out/r2_ric_c2.log 2>&1
LIB=quadruped_ctrl_b200/libquadruped_mpc_b200.so
python tools/ncu_summary.py gpurun_out/r2_ric_c2.ncu-rep smsp__average_warps_issue_stalled sm__pipe_tensor sm__inst_executed_pipe > gpurun_out/r2_ric_c2_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_ric_c2.ncu-rep $LIB riccati 600 > gpurun_out/r2_ric_c2_lines.txt 2>&1
head -3 gpurun_out/r2_ric_c2_lines.txt
rm -f gpurun_out/r2_ric_c2.ncu-rep
