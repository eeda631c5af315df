#!/bin/bash
# Multi-GPU leg (gpurun --gpus N): the 2-GPU peer-gather test, then bench config 2 (weak scaling, NCCL and fused gather)
# and config 4 (65536 problems sharded) at N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus_$N.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k fused_peer 2>&1 | tail -3 | tee gpurun_out/pytest_peer_$N.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
: > gpurun_out/bench_multi_$N.json
$RUN bench.py --gpus $N --steps 100 --warmup 5 2> gpurun_out/m_c2.err | tail -1 | tee -a gpurun_out/bench_multi_$N.json | cut -c1-300
$RUN bench.py --gpus $N --steps 100 --warmup 5 --gather nccl 2> gpurun_out/m_c2p.err | tail -1 | tee -a gpurun_out/bench_multi_$N.json | cut -c1-300
$RUN bench.py --gpus $N --config 4 --steps 30 --warmup 5 2> gpurun_out/m_c4.err | tail -1 | tee -a gpurun_out/bench_multi_$N.json | cut -c1-300
python tools/show_bench.py gpurun_out/bench_multi_$N.json
tail -q -n 2 gpurun_out/m_c2.err gpurun_out/m_c2p.err gpurun_out/m_c4.err
