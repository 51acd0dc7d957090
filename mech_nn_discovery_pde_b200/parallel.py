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


class OverlappedGradReducer:
    """Sum of the learned-parameter gradients over ranks, overlapped with the backward pass (SURVEY.md section 8(f),
    row f3; the reference trains under DDP).  Parameters are grouped into buckets in REVERSE registration order --
    the order in which autograd finishes them; a bucket's flat all-reduce is launched asynchronously from the
    post-accumulate-grad hook of its last parameter, so it runs on the communication stream while the rest of the
    backward (for the discovery models: the PDE layer's adjoint solve and the networks in front of it) is still
    computing.  ``finish()`` waits for the outstanding collectives and scatters the sums back into ``.grad``.

        red = OverlappedGradReducer(model.parameters())
        loss.backward(); red.finish(); optimizer.step()

    Parameters that received no gradient in a step are treated as zero (every rank must launch the same collectives).
    """

    def __init__(self, params, bucket_bytes=25 << 20, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets = []          # lists of parameter indices, in reverse order
        cur, size = [], 0
        for i in reversed(range(len(self.params))):
            p = self.params[i]
            cur.append(i)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        for b, idxs in enumerate(self.buckets):
            for i in idxs:
                self.bucket_of[i] = b
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _active(self):
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, i):
        def hook(_param):
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if not self._active() or self._work[b] is not None:
            return
        ps = [self.params[i] for i in self.buckets[b]]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
        self._flat[b] = flat
        self._work[b] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for the collectives of this step, write the summed gradients back, and re-arm for the next step."""
        for b in range(len(self.buckets)):
            if self._active() and self._work[b] is None:
                self._launch(b)       # a bucket with a parameter that got no gradient this step
            if self._work[b] is not None:
                self._work[b].wait()
                off = 0
                for i in self.buckets[b]:
                    p = self.params[i]
                    n = p.numel()
                    if p.grad is None:
                        p.grad = torch.zeros_like(p)
                    p.grad.copy_(self._flat[b][off:off + n].view_as(p))
                    off += n
            self._work[b] = None
            self._flat[b] = None
            self._pending[b] = len(self.buckets[b])

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
