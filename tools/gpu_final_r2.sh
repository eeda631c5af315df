#!/bin/bash
# Consolidated GPU leg with the Riccati solver as default: GPU tests, smoke, one bench line per BASELINE config, the
# inverse-solver line of config 2 beside it, the reference arm, then the ncu profiles of the Riccati kernel
# (launch lists of configs 2 / 5; full captures of the nv <= 60 and nv <= 96 classes with text summaries made on the box).
set -u
mkdir -p gpurun_out
bash tools/gpu_round2.sh
python bench.py --config 2 --solver inverse --steps 50 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c2_inverse.json
python tools/show_bench.py gpurun_out/bench_c2_inverse.json | grep -v parity
B="python bench.py --steps 3 --warmup 3 --no-cpu"
NCU="ncu --clock-control none"
for c in 2 5; do
  $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/r2_ric_launches_c$c.csv $B --config $c > gpurun_out/r2_ric_launches_c$c.log 2>&1
done
LIB=quadruped_ctrl_b200/libquadruped_mpc_b200.so
for c in 2 5; do
  # the class kernels of a solve are launched back to back; three consecutive Riccati launches contain the non-empty one
  $NCU --set full --import-source on --kernel-name-base demangled -k regex:riccati -s 9 -c 3 -f -o gpurun_out/r2_ric_c$c $B --config $c > gpurun_out/r2_ric_c$c.log 2>&1
  python tools/ncu_summary.py gpurun_out/r2_ric_c$c.ncu-rep smsp__average_warps_issue_stalled sm__pipe_tensor sm__inst_executed_pipe > gpurun_out/r2_ric_c${c}_summary.txt 2>&1
  python tools/ncu_lines.py gpurun_out/r2_ric_c$c.ncu-rep $LIB riccati_kernelILb0 700 > gpurun_out/r2_ric_c${c}_lines.txt 2>&1
  rm -f gpurun_out/r2_ric_c$c.ncu-rep
done
ls -la gpurun_out | head -50
