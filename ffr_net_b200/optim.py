"""Fused clip_grad_value_ + Adam (one kernel launch for all 76 RecNet tensors) — replaces
`clip_grad_value_(recnet.parameters(), 1.0); optim.Adam.step()` of models/trainer.py:185-187.
A torch.optim.Optimizer subclass so LR schedulers and state_dict() keep working: state_dict() has torch.optim.Adam's
layout (per-parameter 'step', 'exp_avg', 'exp_avg_sq'), and load_state_dict() restores the moments and the step count
into the device-resident buffers the kernel uses."""
import numpy as np
import torch

from . import _lib

_CHUNK = 4096


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip_value=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_value=clip_value))
        self._tables = {}          # group index -> (key, table, chunks, n_chunks, hyper)
        self._lr_on_device = {}    # group index -> learning rate last pushed to hyper[0] (NOT part of param_groups)

    def _table(self, gi, group):
        params = [p for p in group["params"] if p.grad is not None]
        for p in params:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in params)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1:]
        rows, chunks = [], []
        for t, p in enumerate(params):
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                raise RuntimeError("FusedClipAdam needs contiguous fp32 CUDA parameters and gradients")
            st = self.state[p]
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            chunks += [(t, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        dev = params[0].device
        table = torch.from_numpy(np.array(rows, dtype=np.int64)).to(dev)
        chunk_t = torch.from_numpy(np.array(chunks, dtype=np.int32)).to(dev)
        if cached is not None:
            hyper = cached[4]                                         # keep lr / step count (and their address)
        else:
            hyper = torch.zeros(2, dtype=torch.float32, device=dev)   # [lr, step] (device-resident: graph-capturable)
            steps = [float(self.state[p]["step"]) for p in params if "step" in self.state[p]]
            if steps:                                                 # resumed from a state_dict
                hyper[1:2].fill_(max(steps))
        self._tables[gi] = (key, table, chunk_t, len(chunks), hyper)
        return table, chunk_t, len(chunks), hyper

    def sync_lr(self):
        """Push the (scheduler-updated) learning rates to the device; call outside graph capture/replay."""
        for gi, group in enumerate(self.param_groups):
            t = self._tables.get(gi)
            if t is not None and self._lr_on_device.get(gi) != group["lr"]:
                t[4][0:1].fill_(float(group["lr"]))
                self._lr_on_device[gi] = group["lr"]

    def device_step(self, gi=0):
        """Adam step count as kept on the device (advances under CUDA-graph replay too). Synchronises."""
        t = self._tables.get(gi)
        return 0 if t is None else int(round(float(t[4][1])))

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            if not any(p.grad is not None for p in group["params"]):
                continue
            table, chunk_t, n_chunks, hyper = self._table(gi, group)
            # lr: pushed to the device only when the scheduler changed it (a host->device copy is not capturable);
            # step: advanced on the device, so a replayed CUDA graph keeps counting
            if self._lr_on_device.get(gi) != group["lr"]:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("learning rate changed while capturing: call step() once before capture")
                hyper[0:1].fill_(float(group["lr"]))
                self._lr_on_device[gi] = group["lr"]
            hyper[1:2].add_(1.0)
            b1, b2 = group["betas"]
            _lib.check(lib.ffr_clip_adam(_lib.ptr(table), _lib.ptr(chunk_t), n_chunks, _lib.ptr(hyper), float(b1),
                                         float(b2), float(group["eps"]), float(group["weight_decay"]),
                                         float(group["clip_value"]), _lib.stream_ptr()), "ffr_clip_adam")
        _lib.bump_weights_generation()       # packed weight caches must be rebuilt
        return None

    # ---- state_dict round trip (torch.optim.Adam layout) -----------------------------------------------------
    def state_dict(self):
        for gi, group in enumerate(self.param_groups):
            step = self.device_step(gi)
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(float(step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        # the state tensors were replaced: rebuild the pointer tables, take lr / step from the loaded state
        old = self._tables
        self._tables = {}
        self._lr_on_device = {}
        for gi, group in enumerate(self.param_groups):
            steps = [float(self.state[p]["step"]) for p in group["params"] if p in self.state and "step" in self.state[p]]
            if gi in old:                     # keep the hyper buffer's address (a captured graph may reference it)
                hyper = old[gi][4]
                hyper[0:1].fill_(0.0)
                hyper[1:2].fill_(max(steps) if steps else 0.0)
                self._tables[gi] = ((), None, None, 0, hyper)

    # ---- snapshot / restore around CUDA-graph capture warm-up ------------------------------------------------
    def snapshot(self):
        snap = {"state": {p: {k: v.detach().clone() for k, v in st.items() if torch.is_tensor(v)}
                          for p, st in self.state.items()},
                "hyper": {gi: t[4].detach().clone() for gi, t in self._tables.items()},
                "had_state": {p: ("exp_avg" in st) for p, st in self.state.items()}}
        return snap

    @torch.no_grad()
    def restore(self, snap):
        for p, st in self.state.items():
            if p in snap["state"] and snap["had_state"].get(p, False):
                for k, v in snap["state"][p].items():
                    if k in st and torch.is_tensor(st[k]) and st[k].shape == v.shape:
                        st[k].copy_(v)
            else:                              # state created during the warm-up: back to a fresh optimizer
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in st:
                        st[k].zero_()
        for gi, t in self._tables.items():
            if gi in snap["hyper"]:
                t[4].copy_(snap["hyper"][gi])
            else:
                t[4][1:2].fill_(0.0)           # step count back to zero, lr stays (same value)
