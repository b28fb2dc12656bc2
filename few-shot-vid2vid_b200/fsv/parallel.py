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
    """Mean all-reduce of the gradients of ``params``, overlapped with backward and without any pack / unpack copies.

    The gradients LIVE in flat buckets: at construction every parameter's ``.grad`` becomes a view into one of a few contiguous
    buffers (filled in reverse parameter order, i.e. roughly the order backward produces them, <= ``bucket_mb`` MiB each), and
    autograd accumulates into those views in place.  A post-accumulate hook per parameter counts arrivals; when a bucket's last
    gradient has landed its all-reduce is launched asynchronously (the collective runs on the process group's own stream while
    backward keeps going on the compute stream).  ``sync()`` after ``backward()`` launches whatever is left (parameters without a
    gradient this iteration contribute zeros), waits, and the optimizer then reads the averaged values through the same views.
    NCCL averages inside the collective (ReduceOp.AVG); other backends (gloo in the CPU tests) sum and scale the bucket once.

    Protocol per optimizer step:  sync.zero()  ->  sync.arm()  ->  loss.backward()  ->  sync()  ->  optimizer.step().
    ``arm`` matters because the generator's backward also runs through the discriminator: D's buckets must not fire then.
    Everything is capturable into a CUDA graph (no host-device synchronisation; NCCL collectives are graph-capturable)."""

    def __init__(self, params, bucket_mb=32, group=None, overlap=None, stream_fire=None):
        self.params = [p for p in params if p.requires_grad]
        on_gpu = any(p.is_cuda for p in self.params)
        if overlap is None:
            # With the side-stream weight gradients (ops.WGRAD_SIDE_STREAM) the leaves must receive their gradients with .grad ==
            # None (adoption launches nothing; an in-place add on the compute stream would race the side stream), so the buckets are
            # filled by multi-tensor copies instead of in place ("collect" mode).
            from . import ops
            overlap = not (ops.WGRAD_SIDE_STREAM and on_gpu)
        if stream_fire is None:
            # collect mode, refined: a post-accumulate hook per parameter counts arrivals, and a bucket whose last gradient has landed
            # is copied and all-reduced right away ON THE SIDE STREAM, i.e. concurrently with the rest of backward; only the buckets
            # that complete at the very end of the pass are left for sync().  (FSV_SYNC_STREAM=0: everything after backward.)
            stream_fire = (not overlap) and os.environ.get('FSV_SYNC_STREAM', '1') != '0'
        self.group, self.overlap, self.stream_fire = group, overlap, stream_fire and not overlap
        limit = max(1, int(bucket_mb * (1 << 20) // 4))
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):
            if cur and size + p.numel() > limit:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += p.numel()
        if cur:
            self.buckets.append(cur)
        self.flat, self.views, self.bucket_of = [], {}, {}
        for bi, bucket in enumerate(self.buckets):
            flat = torch.zeros(sum(p.numel() for p in bucket), device=bucket[0].device, dtype=bucket[0].dtype)
            off = 0
            for p in bucket:
                v = flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    v.copy_(p.grad)
                if overlap:
                    p.grad = v
                self.views[p] = v
                self.bucket_of[p] = bi
                off += p.numel()
            self.flat.append(flat)
        self.armed = False
        self.count = [0] * len(self.buckets)
        self.fired = [False] * len(self.buckets)
        self.handles = []
        hook = self._on_grad if overlap else (self._on_grad_stream if self.stream_fire else None)
        self._hooks = [p.register_post_accumulate_grad_hook(hook) for p in self.params] if hook is not None else []
        backend = dist.get_backend(group) if dist.is_initialized() else 'none'
        self.native_avg = backend == 'nccl'

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_initialized() else 1

    def zero(self):
        """Replaces optimizer.zero_grad(): one memset per bucket, and the gradient views stay bound to the buckets (overlap mode);
        collect mode: gradients set to None (backward then hands each leaf a fresh tensor, no kernel)."""
        if not self.overlap:
            for p in self.params:
                p.grad = None
            return
        for flat in self.flat:
            flat.zero_()
        for p, v in self.views.items():
            if p.grad is not v:
                p.grad = v

    def arm(self):
        self.armed = True
        self.count = [0] * len(self.buckets)
        self.fired = [False] * len(self.buckets)
        self.handles = []

    def _fire(self, bi):
        self.fired[bi] = True
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
        self.handles.append((bi, dist.all_reduce(self.flat[bi], op=op, group=self.group, async_op=True)))

    def _on_grad(self, p):
        if not self.armed:
            return
        bi = self.bucket_of[p]
        if p.grad is not self.views[p]:          # somebody re-bound .grad (e.g. zero_grad(set_to_none=True)): fold it back in
            self.views[p].copy_(p.grad)
            p.grad = self.views[p]
        self.count[bi] += 1
        if self.count[bi] == len(self.buckets[bi]) and not self.fired[bi]:
            self._fire(bi)

    def _side(self, bucket):
        """(side stream context, compute stream) for a bucket's copy + collective; plain context on CPU tensors."""
        if not bucket[0].is_cuda:
            return None, None
        from . import ops
        cur = torch.cuda.current_stream()
        side = ops.side_gather_stream()  # waits for the compute stream and for the weight-gradient lanes in use
        return side, cur

    def _fire_stream(self, bi):
        """collect-mode bucket: copy the adopted gradients into the flat bucket and all-reduce it, both on the side stream."""
        self.fired[bi] = True
        bucket = self.buckets[bi]
        side, cur = self._side(bucket)
        ctx = torch.cuda.stream(side) if side is not None else _Null()
        with ctx:
            have = [p for p in bucket if p.grad is not None and p.grad is not self.views[p]]
            if side is not None:
                for p in have:
                    p.grad.record_stream(side)
            if have:
                torch._foreach_copy_([self.views[p] for p in have], [p.grad for p in have])
            for p in bucket:
                if p.grad is None:
                    self.views[p].zero_()
            if self.world > 1:
                op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
                self.handles.append((bi, dist.all_reduce(self.flat[bi], op=op, group=self.group, async_op=True)))

    def _on_grad_stream(self, p):
        if not self.armed:
            return
        bi = self.bucket_of[p]
        self.count[bi] += 1
        if self.count[bi] == len(self.buckets[bi]) and not self.fired[bi]:
            self._fire_stream(bi)

    def __call__(self):
        """After backward(): launch the buckets that did not fill up, wait for all of them, finish the mean."""
        if self.stream_fire:
            if not self.armed:
                self.arm()
            for bi in range(len(self.buckets)):
                if not self.fired[bi]:
                    self._fire_stream(bi)
            side = self._side(self.buckets[0])[0]
            for bi, h in self.handles:
                h.wait()                 # the compute stream waits for the collective (which waited for the side stream's copy)
                if not self.native_avg:
                    self.flat[bi].mul_(1.0 / self.world)
            if side is not None:
                torch.cuda.current_stream().wait_stream(side)      # world == 1 / zero-fills: still ordered before the optimizer
            for p in self.params:
                p.grad = self.views[p]
            self.armed = False
            self.handles = []
            return
        if not self.overlap:
            have = [p for p in self.params if p.grad is not None and p.grad is not self.views[p]]
            if have:
                torch._foreach_copy_([self.views[p] for p in have], [p.grad for p in have])     # one multi-tensor kernel per chunk
            for p in self.params:
                if p.grad is None:
                    self.views[p].zero_()
                p.grad = self.views[p]
            self.arm()
        if not self.armed:
            self.arm()
        for p, v in self.views.items():
            if p.grad is not v:
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
        for bi in range(len(self.buckets)):
            if not self.fired[bi]:
                self._fire(bi)
        for bi, h in self.handles:
            h.wait()
            if not self.native_avg:
                self.flat[bi].mul_(1.0 / self.world)
        self.armed = False
        self.handles = []


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def broadcast_state(module, src=0, group=None):
    """Make every rank start from rank ``src``'s parameters and buffers (replica 0 is the original module in the
    reference: sync_batchnorm/replicate.py:24-28)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
