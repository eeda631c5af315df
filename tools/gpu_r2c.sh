#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "uniform_host or tick_entry or overflow or host_entry" 2>&1 | tail -4 | tee gpurun_out/pytest_r2c.log
python tools/tick_probe.py 2>&1 | tee gpurun_out/tick_probe.log
python bench.py --config 2 --steps 100 --warmup 5 --no-cpu 2>gpurun_out/b2.err | tail -1 > gpurun_out/r2c_c2.json
python bench.py --config 2 --steps 100 --warmup 5 --no-cpu --e2e-slots 4 2>>gpurun_out/b2.err | tail -1 > gpurun_out/r2c_c2_s4.json
MPC_NO_HOST_CLASSIFY=1 python bench.py --config 2 --steps 100 --warmup 5 --no-cpu 2>>gpurun_out/b2.err | tail -1 > gpurun_out/r2c_c2_nohc.json
python bench.py --config 4 --steps 30 --warmup 5 --no-cpu 2>>gpurun_out/b2.err | tail -1 > gpurun_out/r2c_c4.json
python tools/show_bench.py gpurun_out/r2c_c2.json gpurun_out/r2c_c2_s4.json gpurun_out/r2c_c2_nohc.json gpurun_out/r2c_c4.json | grep -v parity
tail -5 gpurun_out/b2.err
