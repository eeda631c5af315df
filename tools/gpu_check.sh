#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, smoke, both bench arms, ncu launch list + one full capture.
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpu.txt
nproc | tee gpurun_out/nproc.txt
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
if [ "${1:-}" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mpc_solve -s 8 -c 1 -f -o gpurun_out/prof_solve \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/prof_run.log 2>&1
ls -la gpurun_out
fi
