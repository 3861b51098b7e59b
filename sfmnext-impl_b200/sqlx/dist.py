"""Data-parallel plumbing for the hot path: one process per GPU, the batch sharded by sample, and ONE bucketed
all-reduce (average) of the parameter gradients per step (NCCL over NVLink on GPUs; gloo in the CPU tests).

The reference uses nn.DataParallel (trainer.py:74,93: scatter / replicate / gather / reduce_add to GPU 0 every
iteration, loss on GPU 0).  Every loss term is a per-sample mean followed by a batch mean (trainer.py:532,535,546;
layers.py:280), so equal shards + gradient averaging reproduce the full-batch gradient."""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """Slice dim 0 of every tensor in a dict into `world` equal shards and return shard `rank`."""
    out = {}
    for k, v in batch.items():
        if not torch.is_tensor(v):
            out[k] = v
            continue
        if v.shape[0] % world != 0:
            raise ValueError("batch dimension %d of %r is not divisible by world size %d" % (v.shape[0], k, world))
        per = v.shape[0] // world
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def allreduce_flat_(flat, group=None):
    """The path's one exchange step: average, in place, the flat gradient bucket the kernels wrote into
    (HotPath.grad_flat: every parameter gradient is a view of it, so there is no pack / unpack).  Capturable in a CUDA graph
    with the NCCL backend."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:                                       # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(dist.get_world_size(group))
    return flat


class GradBucket:
    """Flat fp32 bucket for gradients that live in separate tensors (e.g. the rest of the model under plain autograd):
    pack -> all_reduce(AVG or SUM/world) -> unpack.  The hot path itself does not need it (allreduce_flat_)."""

    def __init__(self, tensors, group=None):
        self.sizes = [t.numel() for t in tensors]
        self.flat = torch.empty(sum(self.sizes), device=tensors[0].device, dtype=torch.float32)
        self.group = group

    def allreduce_(self, grads):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return grads
        views = list(self.flat.split(self.sizes))
        torch._foreach_copy_(views, [g.reshape(-1) for g in grads])
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            self.flat.div_(dist.get_world_size(self.group))
        torch._foreach_copy_([g.view(-1) for g in grads], views)
        return grads
