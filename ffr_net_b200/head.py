"""CosFace head + CrossEntropy for the training step: AddMarginProduct (models/recnet.py:238-270) followed by
nn.CrossEntropyLoss (models/trainer.py:173-176), fused so that neither the (N,10575) logits nor the one-hot matrix are
materialised. Kernels: ffr_net_b200/csrc/head_kernels.cu (pack / finish / backward / normalisation Jacobian) and the
tcgen05 GEMM of conv_gemm.cu with the EPI_COSFACE epilogue (forward) and plain / split-K epilogues (backward).

`cosface_ce(weight, v, label)` returns (loss, pred): the mean cross-entropy of s*(cos - m*onehot) (differentiable
w.r.t. v and weight) and the arg-max class of the cosines (what trainer.py:150 derives from `pred_label`).
The public RecNet.forward(input, label) still returns the reference's 7-tuple with both (N,10575) tensors
(recnet_train.add_margin_product); the Trainer mirror uses this fused form.
"""
import torch

from . import _lib

S_DEFAULT, M_DEFAULT = 30.0, 0.40
EPI_OUT_F32_ATOMIC, EPI_OUT_F32 = 0x40, 0x800

import weakref

_wcache = {}      # id(parameter) -> (key, packed, transposed, weakref to the parameter); dead entries are evicted


def _ceil(a, b):
    return (a + b - 1) // b * b


def _packed_classes(lib, weight):
    """hi/lo-split normalised class matrix [c_pad x 1536] and its bf16 transpose [512 x c_pad]; repacked (into the same
    buffers) whenever the weights change (torch version counter, or the fused optimizer's generation counter)."""
    key = (weight.data_ptr(), weight._version, _lib.weights_generation(), str(weight.device))
    held = _wcache.get(id(weight))
    if held is not None and held[3]() is not weight:      # id reused by another tensor
        held = None
    if held is not None and held[0] == key:
        return held[1], held[2]
    classes = weight.shape[0]
    c_pad = _ceil(classes, 256)
    if held is not None and held[1].device == weight.device:
        wp, wt = held[1], held[2]
    else:
        wp = torch.empty(c_pad, 1536, dtype=torch.bfloat16, device=weight.device)
        wt = torch.empty(512, c_pad, dtype=torch.bfloat16, device=weight.device)
    _lib.check(lib.ffr_cosface_pack(_lib.ptr(weight.detach()), classes, c_pad, 1, _lib.ptr(wp), _lib.ptr(wt), c_pad,
                                    _lib.stream_ptr()), "cosface_pack(classes)")
    for k in [k for k, v in _wcache.items() if v[3]() is None]:
        del _wcache[k]
    _wcache[id(weight)] = (key, wp, wt, weakref.ref(weight))
    return wp, wt


class _CosFaceCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, weight, label, s, m):
        lib = _lib.load()
        P, st = _lib.ptr, _lib.stream_ptr()
        dev = v.device
        n, classes = v.shape[0], weight.shape[0]
        c_pad, n_pad = _ceil(classes, 256), _ceil(n, 64)
        v = v.detach().contiguous().float()
        w = weight.detach()
        if not w.is_contiguous() or w.dtype != torch.float32 or v.shape[1] != 512 or w.shape[1] != 512:
            raise ValueError("cosface_ce expects fp32 contiguous weight (classes,512) and v (N,512)")
        wp, wt = _packed_classes(lib, weight)
        lab = label.to(torch.int32).contiguous()
        vp = torch.empty(n_pad, 1536, dtype=torch.bfloat16, device=dev)
        vt = torch.empty(512, n_pad, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.ffr_cosface_pack(P(v), n, n_pad, 0, P(vp), P(vt), n_pad, st), "cosface_pack(samples)")
        cos = torch.empty(n, c_pad, dtype=torch.float32, device=dev)
        sumexp = torch.empty(n, dtype=torch.float32, device=dev)
        zlabel = torch.empty(n, dtype=torch.float32, device=dev)
        argkey = torch.empty(n, dtype=torch.int64, device=dev)
        part = torch.empty(n, c_pad // 128, dtype=torch.float32, device=dev)
        _lib.check(lib.ffr_cosface_ce_fwd(P(vp), n, P(wp), c_pad, classes, P(lab), s, m, P(cos), P(sumexp), P(zlabel),
                                          P(argkey), P(part), st), "cosface_ce_fwd")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pred = torch.empty(n, dtype=torch.int64, device=dev)
        _lib.check(lib.ffr_cosface_ce_finish(P(sumexp), P(zlabel), P(argkey), n, s, P(loss), P(pred), st),
                   "cosface_ce_finish")
        ctx.save_for_backward(v, w, cos, sumexp, lab, vt, wt)
        ctx.sm = (s, m, classes, c_pad, n_pad)
        ctx.mark_non_differentiable(pred)
        return loss, pred

    @staticmethod
    def backward(ctx, gloss, _gpred):
        lib = _lib.load()
        P, st = _lib.ptr, _lib.stream_ptr()
        v, w, cos, sumexp, lab, vt, wt = ctx.saved_tensors
        s, m, classes, c_pad, n_pad = ctx.sm
        n, dev = v.shape[0], v.device
        g = gloss.detach().float().contiguous()
        dcos = torch.empty(n, c_pad, dtype=torch.bfloat16, device=dev)
        dcos_t = torch.empty(c_pad, n_pad, dtype=torch.bfloat16, device=dev)
        _lib.check(lib.ffr_cosface_ce_bwd(P(cos), c_pad, classes, n, n_pad, P(lab), P(sumexp), P(g), s, m, P(dcos),
                                          P(dcos_t), st), "cosface_ce_bwd")
        dv = dw = None
        if ctx.needs_input_grad[0]:
            # dv^ (n x 512) = dcos (n x c_pad) . W^ : contraction over the classes, split-K over the SMs
            dvh = torch.zeros(n, 512, dtype=torch.float32, device=dev)
            m_tiles = (n + 127) // 128
            splits = max(1, 148 // (2 * m_tiles))
            _lib.check(lib.ffr_conv_gemm(P(dcos), n, c_pad, c_pad, P(wt), c_pad, 512, 1, None, None, n, 0, 0, 0, 0, 0,
                                         EPI_OUT_F32_ATOMIC, None, None, None, 0, 0, None, P(dvh), None, 0, None,
                                         splits, None, 0, 0, 0, st), "cosface dv GEMM")
            dv = torch.empty_like(v)
            _lib.check(lib.ffr_normalize_bwd(P(v), P(dvh), n, P(dv), st), "normalize_bwd(v)")
        if ctx.needs_input_grad[1]:
            # dW^ (classes x 512) = dcos^T (c_pad x n_pad) . v^ : contraction over the samples
            dwh = torch.empty(classes, 512, dtype=torch.float32, device=dev)
            _lib.check(lib.ffr_conv_gemm(P(dcos_t), c_pad, n_pad, n_pad, P(vt), n_pad, 512, 1, None, None, classes, 0, 0,
                                         0, 0, 0, EPI_OUT_F32, None, None, None, 0, 0, None, P(dwh), None, 0, None, 1,
                                         None, 0, 0, 0, st), "cosface dW GEMM")
            dw = torch.empty_like(w)
            _lib.check(lib.ffr_normalize_bwd(P(w), P(dwh), classes, P(dw), st), "normalize_bwd(W)")
        return dv, dw, None, None, None


def cosface_ce(weight, v, label, s=S_DEFAULT, m=M_DEFAULT):
    """(mean CE loss of the CosFace logits, arg-max class of the cosines). v: (N,512) fp32 CUDA, weight: (C,512)."""
    if not v.is_cuda:
        raise RuntimeError("ffr_net_b200.head.cosface_ce runs only on CUDA (sm_100a); there is no CPU fallback")
    return _CosFaceCE.apply(v, weight, label, float(s), float(m))


class FusedCE:
    """What RecNet's training forward puts in the `pred_loss` / `pred_label` slots of the 7-tuple when the fused head is
    requested (Trainer): the CE loss of that call and the predicted classes, instead of two (N,10575) tensors."""

    def __init__(self, loss, pred):
        self.loss = loss
        self.pred = pred


class GroupedHead:
    """CosFace head + cross-entropy for the batched training step: the pooled features of the G = 2 RecNet calls of an
    iteration (unmasked rows first, then masked; same labels) go through ONE cosine GEMM; each call keeps its own mean
    CE loss (models/trainer.py:173-176). No autograd: forward() fills the per-call losses and predictions, backward()
    returns the gradient w.r.t. the pooled features and writes the classifier gradient. All buffers are allocated once."""

    def __init__(self, classifier, n_total, n_per_group, dev):
        self.cls = classifier
        self.NT, self.n = n_total, n_per_group
        classes = classifier.weight.shape[0]
        self.classes, self.c_pad, self.n_pad = classes, _ceil(classes, 256), _ceil(n_total, 64)
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.vp = torch.zeros(self.n_pad, 1536, **bf)
        self.vt = torch.zeros(512, self.n_pad, **bf)
        self.cos = torch.zeros(n_total, self.c_pad, **f32)
        self.sumexp = torch.zeros(n_total, **f32)
        self.sumexp_part = torch.zeros(n_total, self.c_pad // 128, **f32)
        self.zlabel = torch.zeros(n_total, **f32)
        self.argkey = torch.zeros(n_total, dtype=torch.int64, device=dev)
        self.pred = torch.zeros(n_total, dtype=torch.int64, device=dev)
        self.lab = torch.zeros(n_total, dtype=torch.int32, device=dev)
        self.dcos = torch.zeros(n_total, self.c_pad, **bf)
        self.dcos_t = torch.zeros(self.c_pad, self.n_pad, **bf)
        self.dvh = torch.zeros(n_total, 512, **f32)
        self.dv = torch.zeros(n_total, 512, **f32)
        self.dwh = torch.zeros(classes, 512, **f32)

    def forward(self, v, label, ce_out):
        """v: (G*n,512) fp32; label: (n,) int; ce_out: (G,) fp32 receives the mean CE loss of each call."""
        lib = _lib.load()
        P, st = _lib.ptr, _lib.stream_ptr()
        w = self.cls.weight
        if w.dtype != torch.float32 or not w.is_contiguous():
            raise ValueError("classifier weight must be contiguous fp32")
        self.v = v
        G = self.NT // self.n
        for g in range(G):
            self.lab[g * self.n:(g + 1) * self.n].copy_(label)
        self.wp, self.wt = _packed_classes(lib, w)
        s, m = float(self.cls.s), float(self.cls.m)
        _lib.check(lib.ffr_cosface_pack(P(v), self.NT, self.n_pad, 0, P(self.vp), P(self.vt), self.n_pad, st),
                   "cosface_pack(samples)")
        _lib.check(lib.ffr_cosface_ce_fwd(P(self.vp), self.NT, P(self.wp), self.c_pad, self.classes, P(self.lab), s, m,
                                          P(self.cos), P(self.sumexp), P(self.zlabel), P(self.argkey), P(self.sumexp_part),
                                          st), "cosface_ce_fwd")
        for g in range(G):
            lo = g * self.n
            _lib.check(lib.ffr_cosface_ce_finish(P(self.sumexp[lo:]), P(self.zlabel[lo:]), P(self.argkey[lo:]), self.n, s,
                                                 P(ce_out[g:]), P(self.pred[lo:]), st), "cosface_ce_finish")

    def backward(self, gloss, dw_out, accumulate=False):
        """gloss: (G,) fp32 device tensor, d(total loss)/d(CE of call g). Writes d/d classifier.weight to dw_out
        (overwritten) and returns d/dv (G*n,512)."""
        lib = _lib.load()
        P, st = _lib.ptr, _lib.stream_ptr()
        s, m = float(self.cls.s), float(self.cls.m)
        _lib.check(lib.ffr_cosface_ce_bwd_grouped(P(self.cos), self.c_pad, self.classes, self.NT, self.n_pad, P(self.lab),
                                                  P(self.sumexp), P(gloss), self.n, s, m, P(self.dcos), P(self.dcos_t), st),
                   "cosface_ce_bwd")
        # dv^ (NT x 512) = dcos (NT x c_pad) . W^ ; one CTA per tile, no split: deterministic
        _lib.check(lib.ffr_conv_gemm(P(self.dcos), self.NT, self.c_pad, self.c_pad, P(self.wt), self.c_pad, 512, 1, None,
                                     None, self.NT, 0, 0, 0, 0, 0, EPI_OUT_F32, None, None, None, 0, 0, None, P(self.dvh),
                                     None, 0, None, 1, None, 0, 0, 0, st), "cosface dv GEMM")
        _lib.check(lib.ffr_normalize_bwd(P(self.v), P(self.dvh), self.NT, P(self.dv), st), "normalize_bwd(v)")
        # dW^ (classes x 512) = dcos^T (c_pad x n_pad) . v^ : contraction over the samples of both calls
        _lib.check(lib.ffr_conv_gemm(P(self.dcos_t), self.c_pad, self.n_pad, self.n_pad, P(self.vt), self.n_pad, 512, 1,
                                     None, None, self.classes, 0, 0, 0, 0, 0, EPI_OUT_F32, None, None, None, 0, 0, None,
                                     P(self.dwh), None, 0, None, 1, None, 0, 0, 0, st), "cosface dW GEMM")
        if accumulate:
            raise NotImplementedError("GroupedHead overwrites the classifier gradient")
        _lib.check(lib.ffr_normalize_bwd(P(self.cls.weight), P(self.dwh), self.classes, P(dw_out), st), "normalize_bwd(W)")
        return self.dv
