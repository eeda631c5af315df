#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
P='import sys,json; d=json.loads(sys.stdin.read()); print("value %.2f M/s  ms/step %.4f  host queue ms/step %.4f  e2e %.2f  gather_ok %s results_ok %s" % (d["value"]/1e6, d["ms_per_step"], d["host_queue_ms_per_step"], d["e2e"]["value"]/1e6, d.get("gather_ok"), d.get("results_ok")))'
for v in "--pushbufs 1" "--pushbufs 2"; do
echo "== N=$N $v"; $RUN bench.py --gpus $N --steps 200 --warmup 8 --no-cpu $v 2> gpurun_out/md.err | tail -1 | python -c "$P"; tail -n 2 gpurun_out/md.err | grep -v OMP
done
