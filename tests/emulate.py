"""Torch emulation of the device algorithms (TEST ONLY): same layouts, same packed weights, same tap tables as the
CUDA kernels, computed with plain tensor ops. Used by the CPU tests to validate the host-side packing / folding /
layout logic without a GPU, and by GPU tests as a per-kernel expected value. Never imported by the product."""
import torch


def taps_3x3_flat(G):
    return [((r - 1) * G + (s - 1), 0) for r in range(3) for s in range(3)]


def taps_3x3_s2d(G, C):
    out = []
    for r in range(3):
        for s in range(3):
            dh, ph = (-1, 1) if r == 0 else ((0, 0) if r == 1 else (0, 1))
            dw, pw = (-1, 1) if s == 0 else ((0, 0) if s == 1 else (0, 1))
            out.append((dh * G + dw, (ph * 2 + pw) * C))
    return out


def shifted_rows(a, shift):
    """rows m -> a[m + shift], zero outside [0, rows) (TMA out-of-bounds fill)."""
    out = torch.zeros_like(a)
    n = a.shape[0]
    if shift >= 0:
        if shift < n:
            out[: n - shift] = a[shift:]
    else:
        if -shift < n:
            out[-shift:] = a[: n + shift]
    return out


def conv_gemm(a, wp, cin, taps):
    """D[m,co] = sum_t sum_c A[m+shift_t, choff_t + c] * Wp[co, t*cin + c] in fp32 on bf16-valued operands."""
    a = a.float()
    wp = wp.float()
    acc = torch.zeros(a.shape[0], wp.shape[0], dtype=torch.float32, device=a.device)
    for t, (shift, choff) in enumerate(taps):
        acc += shifted_rows(a[:, choff:choff + cin], shift) @ wp[:, t * cin:(t + 1) * cin].t()
    return acc


def geom(n_img, S, device):
    """valid mask and border class per flat row."""
    G = S + 1
    h = torch.arange(G, device=device).view(G, 1).expand(G, G)
    w = torch.arange(G, device=device).view(1, G).expand(G, G)
    valid = (h < S) & (w < S)
    ch = torch.where(h == 0, 0, torch.where(h == S - 1, 2, 1))
    cw = torch.where(w == 0, 0, torch.where(w == S - 1, 2, 1))
    cls = ch * 3 + cw
    return valid.reshape(-1).repeat(n_img), cls.reshape(-1).repeat(n_img)


def bf16(x):
    return x.to(torch.bfloat16)


def backbone_forward(pk_units, stem, head, bn, x):
    """Emulates ffr_net_b200.Backbone.forward_internal with the packed weights. x fp32 NCHW."""
    import torch.nn.functional as F
    from ffr_net_b200 import layout
    stem_w, stem_b, stem_a = stem
    n = x.shape[0]
    S = x.shape[2]
    w = stem_w.t().reshape(64, 3, 3, 3)
    h = F.conv2d(bf16(x).float(), bf16(w).float(), padding=1) + stem_b.view(1, -1, 1, 1)
    h = torch.where(h > 0, h, h * stem_a.view(1, -1, 1, 1))
    cur = layout.to_flat(h)
    for u in pk_units:
        G = S + 1
        valid, cls = geom(n, S, x.device)
        acc = conv_gemm(cur, u.w1, u.cin, taps_3x3_flat(G))
        acc = acc + u.bias9[cls]
        acc = torch.where(acc > 0, acc, acc * u.slope.view(1, -1))
        acc = acc * valid.view(-1, 1)
        so = S // u.stride
        if u.stride == 2:
            t_nchw = layout.from_flat(bf16(acc), n, S, u.depth)
            t = layout.to_s2d(t_nchw)
            taps = taps_3x3_s2d(so + 1, u.depth)
        else:
            t = bf16(acc)
            taps = taps_3x3_flat(so + 1)
        valid_o, _ = geom(n, so, x.device)
        acc2 = (conv_gemm(t, u.w2, u.depth, taps) + u.b2.view(1, -1)) * valid_o.view(-1, 1)
        pool = acc2.reshape(n, -1, u.depth).sum(1)
        uu = bf16(acc2)
        if u.cin == u.depth:
            sc = layout.to_flat(layout.from_flat(cur, n, S, u.cin)[:, :, ::u.stride, ::u.stride])
        else:
            xs = layout.to_flat(layout.from_flat(cur, n, S, u.cin)[:, :, ::2, ::2])
            sc = bf16((conv_gemm(xs, u.wsc, u.cin, [(0, 0)]) + u.bsc.view(1, -1)) * valid_o.view(-1, 1))
        mean = pool / float(so * so)
        hid = torch.relu(mean @ u.fc1.t())
        gate = torch.sigmoid(hid @ u.fc2.t())                         # [n, depth]
        rows = (so + 1) * (so + 1)
        y = uu.float().reshape(n, rows, u.depth) * gate.view(n, 1, -1) + sc.float().reshape(n, rows, u.depth)
        cur = bf16(y.reshape(n * rows, u.depth))
        S = so
    bn_scale, bn_shift = bn
    body = layout.from_flat(cur, n, S, 512)
    y = body * bn_scale.view(1, -1, 1, 1) + bn_shift.view(1, -1, 1, 1)
    head_w, head_b = head
    acc = cur.float().reshape(n, -1) @ head_w.float().t() + head_b.view(1, -1)
    f = acc / acc.norm(dim=1, keepdim=True)
    return y, f
