import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from quadruped_ctrl_b200 import engine as E, workloads as W
for name, B in (("config2", 4096), ("config2", 65536), ("four_stance", 4096), ("config5", 4096), ("config3", 4096)):
    h = W.HORIZONS[name]
    rec = torch.from_numpy(W.CONFIGS[name](B)).cuda()
    eng = E.MpcBatch(h, B)
    if B == 4096 and name == "config2": print(eng.classes())
    f, _, st = eng.solve_device(rec)
    torch.cuda.synchronize()
    code = (st.cpu().numpy() & 0xff); its = st.cpu().numpy() >> 8
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): eng.solve_device(rec, forces=f, status=st)
    torch.cuda.synchronize()
    n = 20
    ev0.record()
    for _ in range(n): eng.solve_device(rec, forces=f, status=st)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / n
    print(f"{name} B={B} h={h}: {ms:.3f} ms/solve-batch -> {B/ms*1e3:,.0f} solves/s ; status {np.bincount(code)} iters mean {its.mean():.2f} max {its.max()}")
    eng.close()
