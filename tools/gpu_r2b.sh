#!/bin/bash
# GPU leg of this session: tests, bench config 2 (+ host-classify A/B), then the profiles.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r2b.log
python bench.py --config 2 --steps 100 --warmup 5 2>gpurun_out/b2.err | tail -1 > gpurun_out/r2b_c2.json
MPC_NO_HOST_CLASSIFY=1 python bench.py --config 2 --steps 100 --warmup 5 --no-cpu 2>>gpurun_out/b2.err | tail -1 > gpurun_out/r2b_c2_nohc.json
python bench.py --config 4 --steps 30 --warmup 5 --no-cpu 2>>gpurun_out/b2.err | tail -1 > gpurun_out/r2b_c4.json
python tools/show_bench.py gpurun_out/r2b_c2.json gpurun_out/r2b_c2_nohc.json gpurun_out/r2b_c4.json
tail -5 gpurun_out/b2.err
bash tools/gpu_profiles_r2.sh
