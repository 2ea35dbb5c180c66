"""Fused clip_grad_value_ + Adam (one kernel launch for all 76 RecNet tensors) — replaces
`clip_grad_value_(recnet.parameters(), 1.0); optim.Adam.step()` of models/trainer.py:185-187.
A torch.optim.Optimizer subclass so LR schedulers and state_dict() keep working."""
import numpy as np
import torch

from . import _lib

_CHUNK = 4096


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_value=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_value=clip_value))
        self._tables = {}

    def _table(self, gi, group):
        params = [p for p in group["params"] if p.grad is not None]
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in params)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1:]
        rows, chunks = [], []
        for t, p in enumerate(params):
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError("FusedClipAdam needs contiguous fp32 CUDA parameters and gradients")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            chunks += [(t, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        dev = params[0].device
        table = torch.from_numpy(np.array(rows, dtype=np.int64)).to(dev)
        chunk_t = torch.from_numpy(np.array(chunks, dtype=np.int32)).to(dev)
        hyper = torch.zeros(2, dtype=torch.float32, device=dev)      # [lr, step] (device-resident: graph-capturable)
        if cached is not None:
            hyper.copy_(cached[4])
        self._tables[gi] = (key, table, chunk_t, len(chunks), hyper)
        return table, chunk_t, len(chunks), hyper

    def sync_lr(self):
        """Push the (scheduler-updated) learning rates to the device; call outside graph capture/replay."""
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is not None and group.get("_lr_on_device") != group["lr"]:
                t[4][0:1].fill_(float(group["lr"]))
                group["_lr_on_device"] = group["lr"]

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            table, chunk_t, n_chunks, hyper = self._table(gi, group)
            # lr: pushed to the device only when the scheduler changed it (a host->device copy is not capturable);
            # step: advanced on the device, so a replayed CUDA graph keeps counting
            if group.get("_lr_on_device") != group["lr"]:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("learning rate changed while capturing: call step() once before capture")
                hyper[0:1].fill_(float(group["lr"]))
                group["_lr_on_device"] = group["lr"]
            hyper[1:2].add_(1.0)
            b1, b2 = group["betas"]
            _lib.check(lib.ffr_clip_adam(_lib.ptr(table), _lib.ptr(chunk_t), n_chunks, _lib.ptr(hyper), float(b1),
                                         float(b2), float(group["eps"]), float(group["weight_decay"]),
                                         float(group["clip_value"]), _lib.stream_ptr()), "ffr_clip_adam")
        _lib.bump_weights_generation()       # packed bf16 weight caches must be rebuilt
        return None
