"""Data-parallel plumbing: one process per GPU (torchrun), gradient all-reduce only.

Replaces the reference's single-process nn.DataParallel wrapper / disabled apex DDP
(models/models.py:40-46,79-98; util/distributed.py:15-25).  Video clips shard on the batch
dimension; BatchNorm statistics stay rank-local (the reference's effective behaviour, SURVEY.md
section 2a); the only collective on the data path is one mean all-reduce of the gradients per
optimizer step (two per iteration: D then G), issued on flat buckets over NCCL/NVLink.
Works with any torch.distributed backend (the CPU tests use gloo).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def shard_batch(global_batch, rank, world):
    """DistributedSampler-style even split of a global batch (data/custom_dataset_data_loader.py:19-22)."""
    assert global_batch % world == 0, 'global batch must divide evenly across ranks'
    per = global_batch // world
    return rank * per, (rank + 1) * per


class GradSync:
    """Mean all-reduce of the gradients of ``params`` in flat buckets of at most ``bucket_mb`` MiB.

    Call it after ``backward()`` and before ``optimizer.step()``.  Parameters whose grad is None on this
    rank take part with zeros (all ranks must issue identical collectives)."""

    def __init__(self, params, bucket_mb=256, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets, cur, size = [], [], 0
        limit = bucket_mb * (1 << 20) // 4
        for p in self.params:
            if cur and size + p.numel() > limit:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(cur)
        self._flat = [None] * len(self.buckets)

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def __call__(self):
        if self.world == 1:
            return
        handles = []
        for bi, bucket in enumerate(self.buckets):
            n = sum(p.numel() for p in bucket)
            flat = self._flat[bi]
            if flat is None or flat.device != bucket[0].device:
                flat = self._flat[bi] = torch.empty(n, device=bucket[0].device, dtype=bucket[0].dtype)
            off = 0
            for p in bucket:
                v = flat[off:off + p.numel()]
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad.reshape(-1))
                off += p.numel()
            handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        inv = 1.0 / self.world
        for bi, bucket in enumerate(self.buckets):
            handles[bi].wait()
            flat = self._flat[bi]
            off = 0
            for p in bucket:
                g = flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    p.grad = (g * inv).clone()
                else:
                    p.grad.copy_(g).mul_(inv)
                off += p.numel()


def broadcast_state(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers (replica 0 is the original module in the
    reference: sync_batchnorm/replicate.py:24-28)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
