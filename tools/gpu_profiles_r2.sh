#!/bin/bash
# Runs on the GPU box (via gpurun, ONE GPU): ncu launch lists and one `--set full` capture per kernel that matters
# (VERDICT r1 "Next" 9): the nv<=60 kernel on config 2, the nv<=96 kernel on config 5 (FMA sweep and DMMA sweep), the
# nv<=128 and the wrench-space kernels on config 3, and the classify kernel.  Reports land in gpurun_out/ and are
# summarised into profiles/ here afterwards (tools/ncu_summary.py).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu"
NCU="ncu --clock-control none"
for c in 2 3 5; do
  $NCU --metrics gpu__time_duration.sum -c 60 --csv --log-file gpurun_out/r2_launches_c$c.csv $B --config $c > gpurun_out/r2_launches_c$c.log 2>&1
done
full() {  # name, kernel regex, skip, bench args...
  local name=$1 k=$2 skip=$3
  shift 3
  $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$k" -s "$skip" -c 1 -f -o gpurun_out/r2_$name $B "$@" > gpurun_out/r2_$name.log 2>&1
  tail -2 gpurun_out/r2_$name.log
}
full c2_nv60 'mpc_solve_pipe_kernel<(\(int\))?128,' 8 --config 2
full c5_nv96_fma 'mpc_solve_pipe_kernel<(\(int\))?256, (\(int\))?16, (\(int\))?6,' 4 --config 5
full c5_nv96_mma 'mpc_solve_pipe_kernel<(\(int\))?256, (\(int\))?16, (\(int\))?6,' 4 --config 5 --sweep mma
full c3_nv128 'mpc_solve_pipe_kernel<(\(int\))?256, (\(int\))?16, (\(int\))?8,' 4 --config 3
full c3_wrench 'mpc_solve_wrench_kernel' 8 --config 3
full c2_classify 'mpc_classify_kernel' 8 --config 2
ls -la gpurun_out | grep r2_
