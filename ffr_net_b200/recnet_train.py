"""RecNet in training mode (batch-statistics BatchNorm, label head) with autograd — models/recnet.py:398-429 as used
by models/trainer.py:139-187.

What runs where (round 1):
  * the 15 ConvLayers (97 % of RecNet's FLOPs) — forward conv + BN-stat reduction, BN/PReLU/residual apply, and the
    whole backward (gradient fold of the reflection mirrors, PReLU/BN backward, dgrad, wgrad) — are hand-written
    sm_100a kernels behind `torch.autograd.Function`s (libffr_sm100: ffr_conv_gemm, ffr_bn_prelu_fwd/bwd,
    ffr_wgrad3x3, ffr_nchw_to_h9 / ffr_h9_to_nchw);
  * the thin remainder (selfSimilarity, the Conv4Channel MLP, the two per-sample matmuls, the CosFace head and the
    losses) are library ops (ATen/cuBLAS under autograd) for now — fused kernels for them are the next step
    (DESIGN.md §7). Nothing falls back to the CPU; CPU tensors raise.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib, packing

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ASSOCIATIVE_CONV4CHANNEL = True     # see forward_train

_EPI_GEOM, _EPI_STATS, _EPI_PIXMAJOR, _EPI_PIX_DGRAD = 0x8, 0x400, 0x2000, 0x4000
_TAPS9 = (ctypes.c_int * 9)(*[(r - 1) * 9 + (s - 1) for r in range(3) for s in range(3)])
_ZERO9 = (ctypes.c_int * 9)(*([0] * 9))


def _ceil64(c):
    return (c + 63) // 64 * 64


def wgrad_workspace_elems(cout, cin):
    """Per-tap staging elements ffr_wgrad3x3 needs: ceil128(Cout) * ceil256(Cin) (x9 taps)."""
    return ((cout + 127) // 128 * 128) * ((cin + 255) // 256 * 256)


def _pad1(t, n):
    t = t.detach()
    if t.numel() == n and t.dtype == torch.float32 and t.is_contiguous():
        return t                                   # the kernels only read it
    out = torch.zeros(n, dtype=torch.float32, device=t.device)
    out[: t.numel()] = t.float()
    return out


# Running-statistics updates are read-modify-writes of module buffers. When the two RecNet calls of a training step run
# concurrently on two streams (Trainer, opts.two_streams) they are collected here and applied after the join, first
# call first, exactly as the sequential reference does (models/trainer.py:144-145).
_STATS_SINK = None


def _update_running_stats(bn, mean, var, cnt):
    with torch.no_grad():                           # momentum 0.1, unbiased variance (nn.BatchNorm2d in train mode)
        bn.running_mean.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean)
        bn.running_var.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var * (cnt / max(cnt - 1.0, 1.0)))
        bn.num_batches_tracked += 1


class deferred_running_stats:
    """Context manager: ConvLayer forwards inside it record their batch statistics in `self.items` instead of
    updating the BatchNorm buffers; apply() performs the updates in recording order."""

    def __enter__(self):
        global _STATS_SINK
        self.items, self._prev = [], _STATS_SINK
        _STATS_SINK = self.items
        return self

    def __exit__(self, *exc):
        global _STATS_SINK
        _STATS_SINK = self._prev
        return False

    def apply(self):
        for bn, mean, var, cnt in self.items:
            _update_running_stats(bn, mean, var, cnt)


def prepack(model):
    """Pack every ConvLayer weight (forward + dgrad layouts) and the class matrix of the head on the CURRENT stream, so
    that concurrent forward calls only hit the caches."""
    from . import head
    for _, layer in model.conv_layers():
        w = layer.conv2d.weight
        _packed_weights(layer, w, _ceil64(w.shape[1]), _ceil64(w.shape[0]))
    head._packed_classes(_lib.load(), model.classifier.weight)
    model._train_tables(model.classifier.weight.device)


def _packed_weights(layer, weight, cin_p, cout_p):
    """bf16 K-major weights for the forward GEMM and the dgrad GEMM, packed by one kernel and cached until the
    parameter changes (the two RecNet calls of a step and the backward share them)."""
    key = (weight.data_ptr(), weight._version, _lib.weights_generation(), cin_p, cout_p)
    cache = getattr(layer, "_ffr_pack", None)
    if cache is not None and cache[0] == key:
        return cache[1], cache[2]
    lib = _lib.load()
    cout, cin = weight.shape[0], weight.shape[1]
    dev = weight.device
    wp = torch.empty(cout_p, 9 * cin_p, dtype=torch.bfloat16, device=dev)
    wt = torch.empty(cin_p, 9 * cout_p, dtype=torch.bfloat16, device=dev)
    w = weight.detach()
    w = w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()
    _lib.check(lib.ffr_pack_conv3x3(_lib.ptr(w), cout, cin, cout_p, cin_p, _lib.ptr(wp), _lib.ptr(wt),
                                    _lib.stream_ptr()), "ffr_pack_conv3x3")
    layer._ffr_pack = (key, wp, wt)
    return wp, wt


def _conv_gemm(lib, a, wp, cin, cout, m, n_img, flags, out, stats=None, geom=True):
    """3x3 taps on the H9 grid (pitch 9); plain bf16 rows out. Large batches use pixel-major tiles (128 images at one
    pixel, csrc/conv_gemm.cuh EPI_PIXMAJOR): only the 49 interior pixels (forward) / the taps with a non-halo source
    (dgrad, `flags` without geometry) are computed."""
    if lib.ffr_pixmajor_profitable(n_img):
        flags |= _EPI_PIXMAJOR | (0 if (flags & _EPI_GEOM) else _EPI_PIX_DGRAD)
    rc = lib.ffr_conv_gemm(_lib.ptr(a), a.shape[0], a.shape[1], a.stride(0), _lib.ptr(wp), cin, cout, 9, _TAPS9, _ZERO9,
                           m, 81, 9, 7, 1, n_img, flags, None, None, _lib.ptr(out), out.stride(0), 0, None, None, None,
                           0, _lib.ptr(stats), 1, None, 0, 0, 0, _lib.stream_ptr())
    _lib.check(rc, "ffr_conv_gemm")


class _ConvLayerTrain(torch.autograd.Function):
    """ReflectionPad2d(1) -> Conv2d 3x3 -> BatchNorm2d(batch stats) -> PReLU [+ residual] on H9 bf16 rows."""

    @staticmethod
    def forward(ctx, x_h9, weight, gamma, beta, slope, res_h9, layer, tab):
        lib = _lib.load()
        n = x_h9.shape[0] // 81
        cin_p = x_h9.shape[1]
        cout, cin = weight.shape[0], weight.shape[1]
        cout_p = _ceil64(cout)
        dev = x_h9.device
        x_h9 = x_h9.contiguous()
        wp, wt = _packed_weights(layer, weight, cin_p, cout_p)
        z = torch.empty(n * 81, cout_p, dtype=torch.bfloat16, device=dev)
        stats = torch.zeros(2, cout_p, dtype=torch.float32, device=dev)
        _conv_gemm(lib, x_h9, wp, cin_p, cout_p, n * 81, n, _EPI_GEOM | _EPI_STATS, z, stats)
        cnt = float(n * 49)
        mean = stats[0] / cnt
        var = (stats[1] / cnt - mean * mean).clamp_min_(0.0)
        rstd = torch.rsqrt(var + BN_EPS)
        bn = layer.norm.norm
        if _STATS_SINK is not None:                 # concurrent forward calls: the caller applies them in call order
            _STATS_SINK.append((bn, mean[:cout], var[:cout], cnt))
        else:
            _update_running_stats(bn, mean[:cout], var[:cout], cnt)
        g_p, b_p, s_p = _pad1(gamma, cout_p), _pad1(beta, cout_p), _pad1(slope, cout_p)
        out = torch.empty(n * 81, cout_p, dtype=torch.bfloat16, device=dev)
        res = res_h9.contiguous() if res_h9 is not None else None
        _lib.check(lib.ffr_bn_prelu_fwd(_lib.ptr(z), cout_p, _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(g_p), _lib.ptr(b_p),
                                        _lib.ptr(s_p), _lib.ptr(res), cout_p if res is not None else 0, _lib.ptr(out),
                                        cout_p, _lib.ptr(tab), 4, n, cout_p, _lib.stream_ptr()), "ffr_bn_prelu_fwd")
        ctx.save_for_backward(x_h9, weight, z, mean, rstd, g_p, b_p, s_p, tab, wt)
        ctx.has_res = res is not None
        ctx.dims = (n, cin, cin_p, cout, cout_p)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x_h9, weight, z, mean, rstd, g_p, b_p, s_p, tab, wt = ctx.saved_tensors
        n, cin, cin_p, cout, cout_p = ctx.dims
        dev = dout.device
        dout = dout.contiguous()
        dy = torch.empty(n * 81, cout_p, dtype=torch.bfloat16, device=dev)
        dz = torch.empty(n * 81, cout_p, dtype=torch.bfloat16, device=dev)       # halo rows zeroed by the kernel
        dres = torch.empty(n * 81, cout_p, dtype=torch.bfloat16, device=dev) if ctx.has_res else None
        sums = torch.empty(3, cout_p, dtype=torch.float32, device=dev)
        _lib.check(lib.ffr_bn_prelu_bwd(_lib.ptr(dout), cout_p, _lib.ptr(tab), 4, _lib.ptr(z), cout_p, _lib.ptr(mean),
                                        _lib.ptr(rstd), _lib.ptr(g_p), _lib.ptr(b_p), _lib.ptr(s_p), _lib.ptr(dy), cout_p,
                                        _lib.ptr(dres), cout_p, _lib.ptr(sums), _lib.ptr(dz), cout_p, n, cout_p,
                                        _lib.stream_ptr()), "ffr_bn_prelu_bwd")
        dw = torch.empty(cout, cin, 3, 3, dtype=torch.float32, device=dev)
        ws = torch.empty(9 * wgrad_workspace_elems(cout, cin), dtype=torch.float32, device=dev)
        _lib.check(lib.ffr_wgrad3x3(_lib.ptr(dz), cout_p, _lib.ptr(x_h9), cin_p, 0, n, cout, cin, _lib.ptr(dw),
                                    _lib.ptr(ws), _lib.stream_ptr()), "ffr_wgrad3x3")
        dx = None
        if ctx.needs_input_grad[0]:
            # dgrad = the same shifted-row conv with spatially flipped, transposed weights: WT[ci][(8-t)*Cout_p + co]
            dx = torch.empty(n * 81, cin_p, dtype=torch.bfloat16, device=dev)
            _conv_gemm(lib, dz, wt, cout_p, cin_p, n * 81, n, 0, dx)
        return dx, dw, sums[1, :cout], sums[0, :cout], sums[2, :cout], dres, None, None


class _NchwToH9(torch.autograd.Function):
    """fp32 (N,C,7,7) -> bf16 H9 [N*81, cpad] with the reflection halo filled; backward folds the mirrors."""

    @staticmethod
    def forward(ctx, x, cpad):
        lib = _lib.load()
        n, c = x.shape[0], x.shape[1]
        x = x.contiguous().float()
        alloc = torch.empty if cpad == _ceil64(c) else torch.zeros      # the kernel writes channels [0, ceil64(C))
        out = alloc(n * 81, cpad, dtype=torch.bfloat16, device=x.device)
        _lib.check(lib.ffr_nchw_to_h9(_lib.ptr(x), _lib.ptr(out), cpad, 0, n, c, 1, _lib.stream_ptr()), "ffr_nchw_to_h9")
        ctx.dims = (n, c, cpad)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n, c, cpad = ctx.dims
        g = g.contiguous()
        dx = torch.empty(n, c, 7, 7, dtype=torch.float32, device=g.device)
        _lib.check(lib.ffr_h9_to_nchw(_lib.ptr(g), cpad, 0, _lib.ptr(dx), n, c, 1, _lib.stream_ptr()), "ffr_h9_to_nchw")
        return dx, None


class _H9ToNchw(torch.autograd.Function):
    """bf16 H9 [N*81, cpad] -> fp32 (N,C,7,7) (valid pixels); backward writes the gradient with a zero halo."""

    @staticmethod
    def forward(ctx, h9, c):
        lib = _lib.load()
        n, cpad = h9.shape[0] // 81, h9.shape[1]
        h9 = h9.contiguous()
        y = torch.empty(n, c, 7, 7, dtype=torch.float32, device=h9.device)
        _lib.check(lib.ffr_h9_to_nchw(_lib.ptr(h9), cpad, 0, _lib.ptr(y), n, c, 0, _lib.stream_ptr()), "ffr_h9_to_nchw")
        ctx.dims = (n, c, cpad)
        return y

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        n, c, cpad = ctx.dims
        g = g.contiguous().float()
        d = torch.zeros(n * 81, cpad, dtype=torch.bfloat16, device=g.device)
        _lib.check(lib.ffr_nchw_to_h9(_lib.ptr(g), _lib.ptr(d), cpad, 0, n, c, 0, _lib.stream_ptr()), "ffr_nchw_to_h9")
        return d, None


def cosine_sim(x1, x2, dim=1):
    """recnet.py:220-224."""
    x1 = F.normalize(x1, dim=2)
    x2 = F.normalize(x2, dim=2)
    return torch.bmm(x1, x2.permute(0, 2, 1))


def self_similarity(x):
    """selfSimilarity, recnet.py:226-236. Without autograd (the no-grad targets of the loss, trainer.py:157, or any
    inference use) both Grams come from the library's fp32 kernels; when a gradient is required the ATen/cuBLAS ops
    below run under autograd."""
    if not x.is_cuda:
        raise RuntimeError("ffr_net_b200.selfSimilarity runs only on CUDA")
    if not (torch.is_grad_enabled() and x.requires_grad) and tuple(x.shape[1:]) == (512, 7, 7):
        lib = _lib.load()
        xc = x.detach().contiguous().float()
        n = xc.shape[0]
        ss_s = torch.empty(n, 49, 7, 7, dtype=torch.float32, device=x.device)
        ss_c = torch.empty(n, 512, 512, dtype=torch.float32, device=x.device)
        _lib.check(lib.ffr_self_similarity(_lib.ptr(xc), n, _lib.ptr(ss_s), _lib.ptr(ss_c), _lib.stream_ptr()),
                   "ffr_self_similarity")
        return ss_s, ss_c
    h, w = x.size(2), x.size(3)
    v = x.reshape(x.size(0), x.size(1), -1)
    ss_space = cosine_sim(v.permute(0, 2, 1), v.permute(0, 2, 1))
    ss_channel = cosine_sim(v, v)
    return ss_space.reshape(ss_space.size(0), ss_space.size(1), h, w), ss_channel


def self_similarity_space(x):
    """Only the spatial Gram of selfSimilarity (recnet.py:231,234): (N,C,H,W) -> (N,HW,H,W)."""
    h, w = x.size(2), x.size(3)
    v = x.reshape(x.size(0), x.size(1), -1).permute(0, 2, 1)
    ss = cosine_sim(v, v)
    return ss.reshape(ss.size(0), ss.size(1), h, w)


def self_similarity_channel(x):
    """Only the channel Gram of selfSimilarity (recnet.py:232): (N,C,H,W) -> (N,C,C)."""
    v = x.reshape(x.size(0), x.size(1), -1)
    return cosine_sim(v, v)


def add_margin_product(weight, x, label, s=30.0, m=0.40):
    """AddMarginProduct.forward, recnet.py:257-270 (one-hot built on the input's device)."""
    cosine = F.linear(F.normalize(x), F.normalize(weight))
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    output = (one_hot * (cosine - m)) + ((1.0 - one_hot) * cosine)
    return output * s, cosine


def forward_train(model, x, label, fused_ce=False):
    """RecNet.forward for training / label inputs (recnet.py:398-429). fused_ce: the `pred_loss` and `pred_label` slots
    of the 7-tuple hold one head.FusedCE (CE loss + predicted classes, csrc/head_kernels.cu) instead of the two
    (N,10575) tensors — the form the Trainer consumes."""
    if not x.is_cuda:
        raise RuntimeError("ffr_net_b200.RecNet runs only on CUDA (sm_100a); there is no CPU fallback")
    if not model.training:
        raise NotImplementedError("RecNet(label=...) in eval mode is not implemented (the reference only calls the "
                                  "label path while training, models/trainer.py:144-145)")
    pk = model._train_tables(x.device)
    tab = pk[0]
    n = x.shape[0]
    x = x.contiguous().float()

    def conv(layer, h, res=None):
        return _ConvLayerTrain.apply(h, layer.conv2d.weight, layer.norm.norm.weight, layer.norm.norm.bias,
                                     layer.relu.func.weight, res, layer, tab)

    def resblock(blk, h):
        return conv(blk.conv2, conv(blk.conv1, h), res=h)

    ss_space = self_similarity_space(x)                                                  # :399 (spatial half)
    flat = x.reshape(n, 512, 49)
    s = model.Conv4Space
    h = _NchwToH9.apply(torch.cat((x, ss_space), 1), 576)                                # :401
    h = resblock(s[1], conv(s[0], h))
    h = resblock(s[3], conv(s[2], h))
    h = resblock(s[5], conv(s[4], h))
    m_space = torch.sigmoid(_H9ToNchw.apply(h, 49)).reshape(n, 49, 49)                   # :404-405

    # Conv4Channel on cat(X, ss_channel) (:402,:406). ss_channel = Xh Xh^T (Xh = rows of X normalised over HW) is not
    # materialised: Linear(561->32)(cat(X, Xh Xh^T)) = X W0a^T + Xh (Xh^T W0b^T) + b0 by associativity — the same
    # function of (X, W0, b0), so autograd yields the same gradients; saves three (N,512,512) fp32 round trips.
    c = model.Conv4Channel
    if ASSOCIATIVE_CONV4CHANNEL:
        w0 = c[0].weight
        xh = F.normalize(flat, dim=2)
        t = torch.matmul(xh.transpose(1, 2).contiguous(), w0[:, 49:].t())                # (N,49,32)
        g = torch.matmul(flat, w0[:, :49].t()) + torch.bmm(xh, t) + c[0].bias
    else:                                                                                # literal form (validation)
        g = F.linear(torch.cat((flat, self_similarity_channel(x)), 2), c[0].weight, c[0].bias)
    # Linear(32->512) directly followed by Linear(512->32) (c[2]->c[3], c[5]->c[6]; no activation in between,
    # recnet.py:375-380) is applied as the composed 32x32 map W_b W_a, W_b b_a + b_b: same function of the four
    # parameter tensors (autograd differentiates through the small product), without two (N,512,512) intermediates.
    g = F.prelu(g, c[1].func.weight)
    for a_, b_, p_ in ((2, 3, 4), (5, 6, 7)):
        w_ab = torch.matmul(c[b_].weight, c[a_].weight)                                  # (32, 32)
        b_ab = torch.mv(c[b_].weight, c[a_].bias) + c[b_].bias
        g = F.prelu(F.linear(g, w_ab, b_ab), c[p_].func.weight)
    g = F.linear(g, c[8].weight, c[8].bias)
    m_channel = torch.sigmoid(g)                                                         # :406

    feat_space = torch.matmul(flat, m_space).reshape(n, 512, 7, 7)                       # :409,412
    feat_channel = torch.matmul(m_channel, flat).reshape(n, 512, 7, 7)                   # :410,413
    fm = _NchwToH9.apply(torch.cat((torch.flip(feat_channel, [3]), feat_channel), 1), 1024)   # :416-417
    f = model.ChannelFlipMerge
    fc_h9 = resblock(f[1], conv(f[0], fm))                                               # :418
    feat_channel_out = _H9ToNchw.apply(fc_h9, 512)
    cat_h9 = torch.cat((_NchwToH9.apply(feat_space, 512), fc_h9, _NchwToH9.apply(x, 512)), 1)   # :420
    mg = model.Conv4Merge
    feat_new = _H9ToNchw.apply(resblock(mg[1], conv(mg[0], cat_h9)), 512)                # :421
    feat_new_v = feat_new.mean(dim=(2, 3))                                               # :423
    if label is None:
        return feat_new_v, feat_new
    if fused_ce:
        from . import head
        fused = head.FusedCE(*head.cosface_ce(model.classifier.weight, feat_new_v, label, model.classifier.s,
                                              model.classifier.m))
        return feat_new_v, fused, fused, m_space, m_channel, feat_space, feat_channel_out
    pred_loss, pred_label = add_margin_product(model.classifier.weight, feat_new_v, label)   # :428
    return feat_new_v, pred_loss, pred_label, m_space, m_channel, feat_space, feat_channel_out
