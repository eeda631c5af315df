#!/bin/bash
# A/B runs of experimental builds of the engine library on the GPU box (development aid):
#   tools/ab.sh build/exp/lib_a.so build/exp/lib_b.so ...
# For every library: parity spot-check of config 2 against the oracle, quick bench of all workloads, phase clocks.
set -u
mkdir -p gpurun_out
for lib in "$@"; do
  name=$(basename "$lib" .so)
  echo "=== $name" | tee -a gpurun_out/ab.log
  MPC_LIB_PATH=$PWD/$lib python - <<'EOF' 2>&1 | tee -a gpurun_out/ab.log
import numpy as np, torch, sys
sys.path.insert(0, '.')
from quadruped_ctrl_b200 import engine as E, workloads as W
from oracle import oracle as O
for name, B in (("config2", 256), ("four_stance", 64), ("config5", 64)):
    h = W.HORIZONS[name]
    rec = W.CONFIGS[name](B)
    eng = E.MpcBatch(h, B)
    f, s, st = eng.solve_device(torch.from_numpy(rec).cuda(), want_solution=True)
    torch.cuda.synchronize()
    o = O.solve_batch(rec, h, 64)
    den = np.maximum(np.linalg.norm(o["sol"], axis=1), 1.0)
    e = np.linalg.norm(s.cpu().numpy() - o["sol"], axis=1) / den
    print("parity %s: max rel err vs oracle64 %.2e, status %s" % (name, e.max(), np.bincount(st.cpu().numpy() & 0xff)))
    eng.close()
EOF
  MPC_LIB_PATH=$PWD/$lib python tests/quick_bench.py 2>&1 | tee -a gpurun_out/ab.log
  MPC_LIB_PATH=$PWD/$lib python tools/phase_clocks.py config2 4096 2>&1 | tee -a gpurun_out/ab.log
done
