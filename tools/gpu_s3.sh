#!/bin/bash
# GPU leg (one B200): GPU tests, smoke, one bench line per BASELINE config, the reference arm, then the ncu profiles of
# every kernel that matters with their text summaries made on the box (reports over the 64 MiB return limit are dropped).
set -u
mkdir -p gpurun_out
bash tools/gpu_round2.sh
bash tools/gpu_profiles_r2.sh
LIB=quadruped_ctrl_b200/libquadruped_mpc_b200.so
for n in c2_nv60 c5_nv96_fma c5_nv96_mma c3_nv128 c3_wrench c2_classify; do
  r=gpurun_out/r2_$n.ncu-rep
  [ -f $r ] || continue
  python tools/ncu_summary.py $r smsp__average_warps_issue_stalled sm__pipe_tensor sm__inst_executed_pipe smsp__pcsamp > gpurun_out/r2_${n}_summary.txt 2>&1
done
python tools/ncu_regions.py gpurun_out/r2_c2_nv60.ncu-rep $LIB ILi128E 4096 > gpurun_out/r2_c2_nv60_regions.txt 2>&1
python tools/ncu_lines.py gpurun_out/r2_c2_nv60.ncu-rep $LIB ILi128E 40 > gpurun_out/r2_c2_nv60_lines.txt 2>&1
du -sh gpurun_out
# keep the return under the limit: drop the largest reports first
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1)
  [ -n "$big" ] || break
  echo "dropping $big"; rm -f $big
done
ls -la gpurun_out
