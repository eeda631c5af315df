#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, smoke, one bench line per BASELINE config (-> gpurun_out/bench_configs.json),
# the reference arm.  Arguments: "quick" skips the reference arm and uses fewer steps.
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpu.txt
nproc | tee gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
STEPS=${STEPS:-50}
: > gpurun_out/bench_configs.json
for c in 1 2 3 4 5; do
  timeout 900 python bench.py --config $c --steps $STEPS --warmup 5 2> gpurun_out/bench_c$c.err | tail -1 | tee -a gpurun_out/bench_configs.json
  tail -n 3 gpurun_out/bench_c$c.err
done
if [ "${1:-}" != "quick" ]; then
  python bench.py --impl reference --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
fi
