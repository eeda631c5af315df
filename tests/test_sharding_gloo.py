"""N>1 plumbing on the CPU tier: two gloo ranks shard a batch, each solves its range, one all-gather of the
forces, and every rank ends with the unsharded answer.  The per-shard compute is a stand-in here (the
host emulation of the kernel source from tests/emu -- there is no GPU on this tier).  On the GPUs, bench.py cuts
configs 4 and 5 with the same sharding.shard_bounds and gathers with all_gather_into_tensor over NCCL (checked in the
run: `gather_ok`); tests/gpu_peer_gather_check.py covers the fused peer-store gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from common import emu_solve
    from quadruped_ctrl_b200 import workloads as W
    from quadruped_ctrl_b200.sharding import ShardedSolver
    rec = W.config2(total, 10, 77)

    def solve_fn(r):
        return torch.from_numpy(emu_solve(r, 10)["forces"])

    s = ShardedSolver(total, solve_fn)
    gathered = s.solve(s.local_slice(rec))
    if rank == 0:
        q.put((gathered.numpy(), (s.lo, s.hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_two_ranks_gather_the_unsharded_answer(total):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from common import emu_solve
    from quadruped_ctrl_b200 import workloads as W
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, (lo, hi) = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = emu_solve(W.config2(total, 10, 77), 10)["forces"]
    assert gathered.shape == (total, 12)
    assert (gathered == full).all()
    assert (lo, hi) == (0, (total + 1) // 2)
