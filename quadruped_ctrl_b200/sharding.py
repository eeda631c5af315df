"""Batch sharding over GPUs (one process per GPU) and the forces all-gather.

Problems are independent, so a batch of B problems is cut into contiguous, near-equal
ranges, rank r solving [lo_r, hi_r).  The only exchange on the path is the gather of the 12 first-step
forces per problem, issued only when the batch is actually split (north_star).  Backend-agnostic:
NCCL over NVLink on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous near-equal split: the first (total % world) ranks get one extra problem."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(total, world):
    return [shard_bounds(total, world, r)[1] - shard_bounds(total, world, r)[0] for r in range(world)]


def all_gather_forces(local_forces, total, group=None, out=None):
    """local_forces [n_local, 12] -> [total, 12] on every rank, rows in global problem order.

    Equal shards use one all_gather_into_tensor straight into `out`; ragged shards are padded to the
    largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_forces
    sizes = shard_sizes(total, world)
    width = local_forces.shape[1]
    if out is None:
        out = torch.empty((total, width), dtype=local_forces.dtype, device=local_forces.device)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, local_forces.contiguous(), group=group)
        return out
    m = max(sizes)
    padded = torch.zeros((m, width), dtype=local_forces.dtype, device=local_forces.device)
    padded[:local_forces.shape[0]] = local_forces
    buf = torch.empty((world * m, width), dtype=local_forces.dtype, device=local_forces.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    pos = 0
    for r, n in enumerate(sizes):
        out[pos:pos + n] = buf[r * m:r * m + n]
        pos += n
    return out


class ShardedSolver:
    """Solves this rank's contiguous shard of a global batch and gathers everybody's forces.

    solve_fn(records_local) -> forces_local [n_local, 12] (a torch tensor).  On the GPUs it is
    MpcBatch.solve_device; the CPU tests inject a stand-in."""

    def __init__(self, total, solve_fn, group=None):
        self.total = int(total)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(self.total, self.world, self.rank)
        self.solve_fn = solve_fn

    def local_slice(self, records_global):
        return records_global[self.lo:self.hi]

    def solve(self, records_local, out=None):
        forces = self.solve_fn(records_local)
        return all_gather_forces(forces, self.total, self.group, out)
