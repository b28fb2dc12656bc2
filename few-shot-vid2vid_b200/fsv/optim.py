"""Adam on the fsv C ABI: ONE kernel launch updates every parameter of the optimizer (fsv_adam_step), with the step counter on
the device so that the update can be recorded into the training step's CUDA graph.

Drop-in for ``torch.optim.Adam(params, lr, betas)`` as the reference builds it (models/base_model.py:39-48): same update rule
(no amsgrad / weight decay), same ``state_dict`` layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), ``param_groups``
with ``lr`` / ``betas`` / ``eps`` that ``update_learning_rate`` (base_model.py:245-257) can edit between steps.
There is no CPU path: parameters must live on a CUDA device.
"""
import ctypes

import torch

from . import _lib
from ._lib import lib, check, c_vp


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        defaults = dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps))
        super().__init__(params, defaults)
        self._tables = {}          # group index -> dict(key, items_dev, chunks_dev, nchunks, keep)
        self._steps = {}           # group index -> device float scalar (steps taken so far)

    def _state_for(self, gi, p):
        st = self.state[p]
        if gi not in self._steps:
            # a state loaded from a torch.optim.Adam checkpoint carries one 'step' per parameter: adopt its value once
            t0 = float(st['step']) if 'step' in st else 0.0
            self._steps[gi] = torch.full((), t0, device=p.device, dtype=torch.float32)
        if 'exp_avg' not in st:
            st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st['step'] = self._steps[gi]                           # one shared device counter per group (torch keeps one per parameter)
        return st

    N_STAGING = 4      # pinned staging buffers per group: [0] for eager steps, one more for every CUDA-graph capture

    def _table(self, gi, ps):
        """Device table of (param, grad, exp_avg, exp_avg_sq, numel).  Gradients are fresh tensors every eager iteration, so their
        addresses are re-uploaded when they changed (one 40-byte record per parameter from pinned memory, stream-ordered); inside a
        CUDA graph they are static: a table uploaded during capture gets its own device + pinned buffers, which the graph's
        memcpy node keeps reading on every replay."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr()) for p in ps)
        tab = self._tables.get(gi)
        capturing = torch.cuda.is_current_stream_capturing()
        if tab is not None and tab['key'] == key and (tab['captured'] or not capturing):
            return tab
        n = len(ps)
        dev = ps[0].device
        if tab is None or tab['n'] != n or tab['numels'] != [p.numel() for p in ps]:
            if capturing:
                raise _lib.FsvError('fsv Adam: the first step (table allocation) must run before CUDA-graph capture')
            chunk_item, chunk_idx = [], []
            for i, p in enumerate(ps):
                c = int(lib.fsv_adam_chunks(p.numel()))
                chunk_item += [i] * c
                chunk_idx += list(range(c))
            chunks = torch.tensor(list(zip(chunk_item, chunk_idx)), dtype=torch.int32).reshape(-1).to(dev)
            nbytes = ctypes.sizeof(_lib.AdamItem) * n
            tab = dict(n=n, numels=[p.numel() for p in ps], chunks_dev=chunks, nchunks=len(chunk_item), captured=False, key=None,
                       staging=[torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(self.N_STAGING)], used=1,
                       dev_tables=[torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(self.N_STAGING)], slot=0)
            self._tables[gi] = tab
        items = (_lib.AdamItem * n)()
        for i, (p, k) in enumerate(zip(ps, key)):
            if not (p.is_contiguous() and p.grad.is_contiguous() and p.dtype == torch.float32 and p.grad.dtype == torch.float32):
                raise _lib.FsvError('fsv Adam: contiguous float32 parameters and gradients expected')
            items[i].param, items[i].grad, items[i].exp_avg, items[i].exp_avg_sq = k
            items[i].numel = p.numel()
        slot = 0
        if capturing:
            if tab['used'] >= self.N_STAGING:
                raise _lib.FsvError('fsv Adam: more than %d CUDA-graph captures of one optimizer' % (self.N_STAGING - 1))
            slot = tab['used']
            tab['used'] += 1
        if slot == 0 and tab.get('event') is not None:
            tab['event'].synchronize()             # the previous eager step's copy out of this staging buffer is done (certainly, not just likely)
        ctypes.memmove(tab['staging'][slot].data_ptr(), ctypes.addressof(items), ctypes.sizeof(items))
        tab['dev_tables'][slot].copy_(tab['staging'][slot], non_blocking=True)
        if slot == 0:
            tab['event'] = torch.cuda.Event()
            tab['event'].record()
        tab['key'], tab['slot'], tab['captured'] = key, slot, capturing
        tab['items_dev'] = tab['dev_tables'][slot]
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group['params'] if p.grad is not None]
            if not ps:
                continue
            _lib.require_cuda(*ps)
            for p in ps:
                self._state_for(gi, p)
            tab = self._table(gi, ps)
            b1, b2 = group['betas']
            check(lib.fsv_adam_step(c_vp(tab['items_dev'].data_ptr()), c_vp(tab['chunks_dev'].data_ptr()), tab['nchunks'],
                                    c_vp(self._steps[gi].data_ptr()), float(group['lr']), float(b1), float(b2), float(group['eps']),
                                    _lib.stream()), 'fsv_adam_step')
        return loss
