#!/bin/bash
# N-GPU diagnosis of what the per-step gather costs: ms per step of config 2 with no gather, a
# payload-free NCCL gather (the rendezvous alone), the NCCL gather, the fused peer-store epilogue, the copy-engine
# gather with three and with six regions.   gpurun --gpus 2 -- 'bash tools/gpu_multi_diag.sh 2'
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
P='import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f M/s  ms/step %.4f  host queue ms/step %.4f  e2e %.2f  gather_ok %s results_ok %s" % (d["value"]/1e6, d["ms_per_step"], d["host_queue_ms_per_step"], d["e2e"]["value"]/1e6, d.get("gather_ok"), d.get("results_ok")))'
run() { echo "== N=$N $*"; env $ENVV $RUN bench.py --gpus $N --steps 200 --warmup 8 --no-cpu "$@" 2> gpurun_out/md.err | tail -1 | python -c "$P"; }
ENVV="X=0" run --gather none
ENVV="MPC_DIAG_TINY_GATHER=1" run --gather nccl
ENVV="X=0" run --gather nccl
ENVV="X=0" run --gather peer
ENVV="X=0" run --gather push --pushbufs 1
ENVV="X=0" run --gather push --pushbufs 2
