"""Small run of every kernel variant for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from quadruped_ctrl_b200 import engine as E, records as R, ticks as T, workloads as W  # noqa: E402

# run with MPC_RIC_ALWAYS=1 so that these small batches go through the Riccati kernel where it is the class's solver
for solver in ("riccati", "inverse"):
    for name, B in (("config2", 96), ("config5", 48), ("four_stance", 32), ("config3", 48)):
        h = W.HORIZONS[name]
        rec = W.CONFIGS[name](B)
        if name == "four_stance":
            rec.view(np.float32)[:8, R.REC_FMAX] = 6.0   # large working sets: tile overflow -> slab (riccati) / re-queue (inverse)
        eng = E.MpcBatch(h, B)
        eng.set_solver(solver)
        f, s, st = eng.solve_host(rec, want_solution=True)
        print(solver, name, [c["threads"] for c in eng.classes()], np.bincount(st & 0xff), "iters max", (st >> 8).max())
        eng.close()
eng = E.MpcBatch(10, 64)
tk = torch.from_numpy(T.synth_ticks(64, 10, 3, mixed_gaits=True)).cuda()
out = eng.solve_ticks_device(tk, want_solution=True)
torch.cuda.synchronize()
print("ticks", np.bincount(out[2].cpu().numpy() & 0xff))
