"""Multi-GPU plumbing: one process per GPU, the batch of independent PDE instances is sharded
contiguously, and the solve itself never communicates (each rank runs its own Krylov space over its
shard).  The only collective of a discovery training step is the sum of the
learned-parameter gradients, done here with torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(global_batch, rank, world):
    """Contiguous shard [lo, hi) of a batch; earlier ranks take the remainder."""
    base, rem = divmod(int(global_batch), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_batch(t, rank, world, dim=0):
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def allreduce_param_grads(params, group=None):
    """Sum .grad of the learned parameters over ranks with ONE flat all-reduce (ParamNets of the GL model are
    ~50 MB fp64: latency bound, so one bucket)."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
