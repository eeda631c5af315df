#!/bin/bash
# 8-GPU leg (gpurun --gpus 8, charged 8x: kept short): bench config 2 (weak scaling, NCCL gather -- what the driver's
# SCALE run launches) and config 4 (65536 problems sharded over the 8 GPUs).
set -u
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
: > gpurun_out/bench_multi_$N.json
$RUN bench.py --gpus $N --steps 100 --warmup 5 --no-cpu 2> gpurun_out/m8_c2.err | tail -1 | tee -a gpurun_out/bench_multi_$N.json | cut -c1-200
$RUN bench.py --gpus $N --config 4 --steps 30 --warmup 5 --no-cpu 2> gpurun_out/m8_c4.err | tail -1 | tee -a gpurun_out/bench_multi_$N.json | cut -c1-200
python tools/show_bench.py gpurun_out/bench_multi_$N.json | grep -v parity
tail -q -n 2 gpurun_out/m8_c2.err gpurun_out/m8_c4.err
