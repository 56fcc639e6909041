"""Adam for the train step (reference legacy_code/train.py:112-114 builds torch.optim.Adam over both networks and
train.py:55 steps it): same constructor arguments and update rule as torch.optim.Adam (no amsgrad), every parameter
tensor updated by ONE launch of wnb200_adam_step.  State (`exp_avg`, `exp_avg_sq`, fp32) lives in the optimizer's
`state` dict under torch's key names, `state_dict()` / `load_state_dict()` interchange with torch.optim.Adam's."""
import ctypes

import numpy as np
import torch

from . import _lib, ops


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("wavenet_speech_b200.optim.Adam: amsgrad is not implemented")
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._tables = {}          # group index -> (key, items tensor, chunks tensor, nchunks)

    def _table(self, gi, entries, device):
        """Device tables for one group; rebuilt only when a pointer moved (gradients usually come back at the same
        addresses from the caching allocator, flat-bucket gradients always do)."""
        key = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()) for p, g, m, v in entries)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit
        chunk = _lib.load().wnb200_adam_chunk_elems()
        items = (_lib.AdamItem * len(entries))()
        chunks = []
        for i, (p, g, m, v) in enumerate(entries):
            it = items[i]
            it.param, it.grad, it.exp_avg, it.exp_avg_sq = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
            it.numel = p.numel()
            it.is_bf16 = 1 if p.dtype == torch.bfloat16 else 0
            for e0 in range(0, p.numel(), chunk):
                chunks.append((i, e0))
        raw = np.frombuffer(bytes(items), dtype=np.uint8).copy()
        items_d = torch.from_numpy(raw).to(device)                       # pageable copy: complete when this returns
        chunks_d = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 2).to(device)
        hit = (key, items_d, chunks_d, len(chunks))
        self._tables[gi] = hit
        return hit

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                ops._need_cuda(p)
                g = p.grad
                if g.is_sparse or g.dtype != p.dtype or p.dtype not in (torch.float32, torch.bfloat16):
                    raise RuntimeError("wavenet_speech_b200.optim.Adam: dense fp32 / bf16 gradients of the parameter's dtype")
                if not p.is_contiguous():
                    raise RuntimeError("wavenet_speech_b200.optim.Adam: parameters must be contiguous")
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0.0        # a Python number (torch.optim.Adam accepts it in a loaded state_dict and
                                            # turns it into its 0-dim tensor): 450 tensor increments per step are 1 ms of host time
                    st["exp_avg"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                    st["exp_avg_sq"] = torch.zeros(p.shape, dtype=torch.float32, device=p.device)
                entries.append((p, g, st["exp_avg"], st["exp_avg_sq"]))
            if not entries:
                continue
            # one launch per distinct step count (one, unless a parameter sat a step out: the bias correction is per tensor)
            by_t = {}
            for e in entries:
                st = self.state[e[0]]
                t = int(float(st["step"])) + 1
                st["step"] = float(t)
                by_t.setdefault(t, []).append(e)
            dev = entries[0][0].device
            b1, b2 = group["betas"]
            with torch.cuda.device(dev):
                for rank, t in enumerate(sorted(by_t)):
                    _, items_d, chunks_d, nchunks = self._table((gi, rank), by_t[t], dev)
                    _lib.call("wnb200_adam_step", nchunks, ops._p(items_d), ops._p(chunks_d), float(group["lr"]), float(b1),
                              float(b2), float(group["eps"]), float(group["weight_decay"]), t, ops._stream())
        return loss
