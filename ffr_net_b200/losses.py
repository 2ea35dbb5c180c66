"""Similarity / triplet / identity losses of Trainer.backward (models/trainer.py:31-43, :154-171) with their gradients,
as sequences of library calls (csrc/loss_kernels.cu + the tcgen05 GEMM). No autograd, no ATen arithmetic; every buffer is
allocated once per batch shape. Layout conventions: feature maps are fp32 "H9" matrices [n_img*81][512] whose own rows
(pixel (h,w) at row (h+1)*9 + (w+1)) hold the values — what recnet_train.TrainEngine produces and consumes."""
import ctypes

import torch

from . import _lib

EPI = _lib.EPI
_P = _lib.ptr


class LossWorkspace:
    """Buffers for a step of G = 2 calls of n samples each (group 0 = unmasked, group 1 = masked)."""

    def __init__(self, n, dev):
        NT = 2 * n
        self.n, self.NT = n, NT
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.A6 = torch.zeros(NT * 512, 384, **bf)
        self.B6 = torch.zeros(NT * 512, 384, **bf)
        self.FhT = torch.zeros(NT * 64, 512, **bf)
        self.inv_f = torch.zeros(NT * 512, **f32)
        self.D = torch.zeros(NT * 512, 512, **bf)
        self.spart = torch.zeros(NT * 16 * 2 * 512, **f32)
        self.e = torch.zeros(NT * 512, 64, **f32)
        self.dfc = torch.zeros(NT * 81, 512, **f32)      # gradient w.r.t. feat_channel (own rows; halo rows stay zero)
        self.dfs = torch.zeros(NT * 81, 512, **f32)      # gradient w.r.t. feat_space
        self.chan_sums = torch.zeros(2 * 64, **f32)
        self.space_part = torch.zeros(NT, **f32)
        self.row_part = torch.zeros(n * 5, **f32)
        self.dvl = torch.zeros(NT, 512, **f32)           # triplet + identity gradient w.r.t. the pooled features
        self.out = torch.zeros(8, **f32)


def selfsim_channel(lw, fc, x_non, n_img, n, group0, weight0, want_grad=True):
    """MSE(ss_channel(X), ss_channel(feat_channel)) terms of trainer.py:158-165 for n_img = G*n samples whose first
    group is `group0` (0 / 1). fc: fp32 H9 [n_img*81][512]; x_non: (n,512,7,7) targets. Fills lw.chan_sums[group0..]
    (sums of squared Gram differences) and, with want_grad, lw.dfc rows [group0*n*81, ...) with the gradient of
    weight0 * L1 (weight0 = loss_weight[0]; the 1/4 of the two nested means is applied here)."""
    lib = _lib.load()
    st = _lib.stream_ptr()
    G = n_img // n
    _lib.check(lib.ffr_selfsim_channel_pack(_P(fc), 512, _P(x_non), n_img, n, _P(lw.A6), _P(lw.B6), _P(lw.FhT),
                                            _P(lw.inv_f), st), "selfsim_channel_pack")
    d = _lib.ConvGemmDesc()                  # D = F^F^T - X^X^T (bf16 out) + per-tile sums of D^2
    d.a, d.a_rows, d.a_cols, d.a_ld = _P(lw.A6), n_img * 512, 384, 384
    d.wp, d.Cin, d.Cout, d.ntaps = _P(lw.B6), 384, 512, 1
    d.M = n_img * 512
    d.flags = EPI.STATS
    d.out, d.ldo = _P(lw.D), 512
    d.stats_part = _P(lw.spart)
    d.num_splits, d.b_rows_per_mtile, d.b_mtile_div = 1, 512, 4
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "selfsim D GEMM")
    _lib.check(lib.ffr_sumsq_reduce(_P(lw.spart), n * 16, 512, G, _P(lw.chan_sums[group0 * 64:]), st), "sumsq_reduce")
    if not want_grad:
        return
    d = _lib.ConvGemmDesc()                  # e = D F^ (per sample: batched K-major F^T)
    d.a, d.a_rows, d.a_cols, d.a_ld = _P(lw.D), n_img * 512, 512, 512
    d.wp, d.Cin, d.Cout, d.ntaps = _P(lw.FhT), 512, 64, 1
    d.M = n_img * 512
    d.flags = EPI.OUT_F32
    d.out_f32 = _P(lw.e)
    d.num_splits, d.b_rows_per_mtile, d.b_mtile_div = 1, 64, 4
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "selfsim D F GEMM")
    coef = weight0 / (n * 512.0 * 512.0)     # 4 * (weight0 / 4) / (n * 512 * 512), see csrc/loss_kernels.cu
    dfc = lw.dfc[group0 * n * 81:(group0 * n + n_img) * 81]
    _lib.check(lib.ffr_selfsim_channel_bwd(_P(lw.e), _P(fc), 512, _P(lw.inv_f), coef, _P(dfc), 512, n_img, st),
               "selfsim_channel_bwd")


def selfsim_space(lw, fs, x_non, n_img, n, group0, weight0, want_grad=True):
    """MSE(ss_space(X), ss_space(feat_space)) terms (trainer.py:158-164): lw.space_part[group0*n ..] per-sample sums of
    squared differences, lw.dfs the gradient of weight0 * L1."""
    lib = _lib.load()
    st = _lib.stream_ptr()
    coef = weight0 / (n * 2401.0)
    dfs = lw.dfs[group0 * n * 81:(group0 * n + n_img) * 81]
    _lib.check(lib.ffr_selfsim_space_loss(_P(fs), 512, _P(x_non), n_img, n, coef, _P(lw.space_part[group0 * n:]),
                                          _P(dfs) if want_grad else None, 512, st), "selfsim_space_loss")


def triplet_identity(lw, f_non, f_ocl, e_non, e_ocl, w_trip, w_id, margin=0.1):
    """TripletLoss(f_ocl, e_non, e_ocl) (trainer.py:167-169) and the identity MSE (:171): lw.row_part, lw.dvl."""
    lib = _lib.load()
    n = lw.n
    _lib.check(lib.ffr_triplet_identity(_P(f_non), _P(f_ocl), _P(e_non), _P(e_ocl), n, w_trip, w_id, margin,
                                        _P(lw.row_part), _P(lw.dvl[:n]), _P(lw.dvl[n:]), _lib.stream_ptr()),
               "triplet_identity")


def finalize(lw, ce, weights):
    """lw.out[0..3] = weighted loss items (SelfSimilarity, Triplet, Identity, Classifier), [4] / [5] mean pos / neg
    distance, [6] total. ce: (2,) device tensor with the mean CE of the two calls."""
    lib = _lib.load()
    w = [float(v) for v in weights]
    _lib.check(lib.ffr_loss_finalize(_P(lw.space_part), _P(lw.chan_sums), _P(lw.row_part), _P(ce), lw.n, 2,
                                     w[0], w[1], w[2], w[3], _P(lw.out), _lib.stream_ptr()), "loss_finalize")
