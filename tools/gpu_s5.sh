#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu --config 2"
ncu --clock-control none --set full --import-source on --kernel-name-base demangled -k regex:riccati -s 9 -c 3 -f -o gpurun_out/r2_ric_c2 $B > gpurun_out/r2_ric_c2.log 2>&1
tail -2 gpurun_out/r2_ric_c2.log
LIB=quadruped_ctrl_b200/libquadruped_mpc_b200.so
python tools/ncu_summary.py gpurun_out/r2_ric_c2.ncu-rep smsp__average_warps_issue_stalled sm__pipe_tensor sm__inst_executed_pipe > gpurun_out/r2_ric_c2_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_ric_c2.ncu-rep $LIB riccati 600 > gpurun_out/r2_ric_c2_lines.txt 2>&1
head -5 gpurun_out/r2_ric_c2_lines.txt
