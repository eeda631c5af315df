"""Run under torchrun with 2+ GPUs: checks the fused peer-store gather and the copy-engine gather against the NCCL
all-gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_ctrl_b200 import engine as E, workloads as W  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, h = 1024, 10
eng = E.MpcBatch(h, B, local)
rec = torch.from_numpy(W.config2(B, h, 500 + rank)).to(dev)
forces, _, status = eng.solve_device(rec)
ref = torch.empty((world * B, 12), dtype=torch.float32, device=dev)
dist.all_gather_into_tensor(ref, forces)
buf = eng.setup_peer_gather(world * B, rank * B)
dist.barrier()
ok = True
for rep in range(5):  # several epochs: the device-side barrier must order every one of them
    rec_r = torch.from_numpy(W.config2(B, h, 500 + rank + 10 * rep)).to(dev)
    f_r, _, _ = eng.solve_device(rec_r)
    eng.gather_sync()
    snap = buf.clone()          # stream-ordered behind the barrier: must already hold every rank's rows
    ref_r = torch.empty_like(ref)
    dist.all_gather_into_tensor(ref_r, f_r)
    torch.cuda.synchronize()
    ok = ok and bool(torch.equal(snap, ref_r))
    dist.barrier()              # nobody starts the next epoch's stores before everyone has snapshotted
# both scratch slots, two batches in flight: slot q's solve stores into every rank's slot-q region
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for pair in range(3):
    snaps, refs = [], []
    for q in (0, 1):
        rec_r = torch.from_numpy(W.config2(B, h, 900 + rank + 10 * (2 * pair + q))).to(dev)
        torch.cuda.current_stream().synchronize()
        with torch.cuda.stream(streams[q]):
            f_r, _, _ = eng.solve_device(rec_r, stream=streams[q], slot=q)
            eng.gather_sync(stream=streams[q], slot=q)
            snaps.append(eng.gather_views[q].clone())
        refs.append(f_r)
    torch.cuda.synchronize()
    for q in (0, 1):
        ref_r = torch.empty_like(ref)
        dist.all_gather_into_tensor(ref_r, refs[q])
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(snaps[q], ref_r))
    dist.barrier()
# the copy-engine gather: epilogue off, this rank's forces pushed into every rank's region by DMA on a stream of its
# own, then the same flag barrier -- several epochs on two regions, against NCCL
eng.set_gather_fused(False)
comm = torch.cuda.Stream()
ok_push = True
for rep in range(4):
    q = rep & 1
    rec_r = torch.from_numpy(W.config2(B, h, 1300 + rank + 10 * rep)).to(dev)
    torch.cuda.current_stream().synchronize()
    with torch.cuda.stream(streams[q]):
        f_r, _, _ = eng.solve_device(rec_r, stream=streams[q], slot=q)
    comm.wait_stream(streams[q])
    with torch.cuda.stream(comm):
        eng.gather_push(f_r, slot=q, stream=comm)
        snap = eng.gather_views[q].clone()   # stream-ordered behind the barrier
    torch.cuda.synchronize()
    ref_r = torch.empty_like(ref)
    dist.all_gather_into_tensor(ref_r, f_r)
    torch.cuda.synchronize()
    ok_push = ok_push and bool(torch.equal(snap, ref_r))
    dist.barrier()
print("rank %d: peer-store gather %s NCCL all-gather, copy-engine gather %s NCCL all-gather (%d rows, both slots)" %
      (rank, "==" if ok else "!=", "==" if ok_push else "!=", world * B), flush=True)
ok = ok and ok_push
t = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
eng.close()
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
