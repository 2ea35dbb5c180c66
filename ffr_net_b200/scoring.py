"""LFW-style verification scoring — drop-in for the hot functions of `lfw/lfw_eval.py` of the reference
(KFold :110-118, eval_acc :137-153, find_best_threshold :155-162, calculate_distance :226-252,
get_fold_accuracy :255-259, get_accuracy :261-270, get_avg_accuracy :272-287).

The per-pair cosine and the whole 10-fold x 400-threshold sweep run on the device (libffr_sm100.so); the reference
runs the sweep as pure-Python loops fanned out over mp.Pool(10). Decisions are made exactly as the reference makes
them: (double)fp32_score > threshold against the bit-exact np.arange(-1, 1, 0.005) grid, last-best `>=` tie rule.
"""
import numpy as np
import torch

from . import _lib


def thresholds_grid():
    return np.arange(-1.0, 1.0, 0.005)


def KFold(n=6000, n_folds=10, shuffle=False):
    if shuffle:
        raise NotImplementedError("the reference only ever calls KFold(shuffle=False) (lfw_eval.py:274)")
    base = list(range(n))
    folds = []
    for i in range(n_folds):
        test = base[i * n // n_folds:(i + 1) * n // n_folds]
        train = base[:i * n // n_folds] + base[(i + 1) * n // n_folds:]
        folds.append([train, test])
    return folds


def pair_cosine(f1, f2):
    """Row-wise cosine of two (P, D) fp32 CUDA tensors -> (P,) fp32 CUDA tensor (lfw_eval.py:246,248)."""
    if not (f1.is_cuda and f2.is_cuda):
        raise RuntimeError("ffr_net_b200.scoring runs only on CUDA; there is no CPU fallback")
    f1 = f1.contiguous().float()
    f2 = f2.contiguous().float()
    assert f1.shape == f2.shape and f1.dim() == 2
    out = torch.empty(f1.shape[0], dtype=torch.float32, device=f1.device)
    lib = _lib.load()
    _lib.check(lib.ffr_pair_cosine(_lib.ptr(f1), _lib.ptr(f2), _lib.ptr(out), f1.shape[0], f1.shape[1],
                                   _lib.stream_ptr()), "pair_cosine")
    return out


def threshold_sweep(scores, labels, n_folds=10, thresholds=None):
    """Device K-fold sweep. scores (n,) fp32 CUDA, labels (n,) int CUDA. Returns a dict of python lists:
    best_idx, best_thr, test_correct, train_correct, test_acc, and avg_acc (reference divides by the literal 10)."""
    if not scores.is_cuda:
        raise RuntimeError("ffr_net_b200.scoring runs only on CUDA; there is no CPU fallback")
    thr = thresholds_grid() if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    dev = scores.device
    thr_d = torch.from_numpy(thr).to(dev)
    scores = scores.contiguous().float()
    labels = labels.to(device=dev, dtype=torch.int32).contiguous()
    n = scores.shape[0]
    best_idx = torch.empty(n_folds, dtype=torch.int32, device=dev)
    best_thr = torch.empty(n_folds, dtype=torch.float64, device=dev)
    test_c = torch.empty(n_folds, dtype=torch.int32, device=dev)
    train_c = torch.empty(n_folds, dtype=torch.int32, device=dev)
    lib = _lib.load()
    _lib.check(lib.ffr_threshold_sweep(_lib.ptr(scores), _lib.ptr(labels), _lib.ptr(thr_d), n, len(thr), n_folds,
                                       _lib.ptr(best_idx), _lib.ptr(best_thr), _lib.ptr(test_c), _lib.ptr(train_c),
                                       _lib.stream_ptr()), "threshold_sweep")
    bounds = [i * n // n_folds for i in range(n_folds + 1)]
    tc = test_c.tolist()
    out = dict(best_idx=best_idx.tolist(), best_thr=best_thr.tolist(), test_correct=tc,
               train_correct=train_c.tolist(),
               test_acc=[1.0 * tc[f] / (bounds[f + 1] - bounds[f]) for f in range(n_folds)])
    out["avg_acc"] = sum(out["test_acc"]) / (10 if n_folds == 10 else n_folds)
    return out


def calculate_distance(data_loader, encoder, recnet, flag=0, use_flip=False, use_gpu=True):
    """lfw_eval.py:226-252: returns two float64 arrays (n,3) [score, label, idx] for rectified and raw features."""
    d_new, d_raw, lab, idx = [], [], [], []
    for data in data_loader:
        img1, img2 = data["img1"].cuda(non_blocking=True), data["img2"].cuda(non_blocking=True)
        with torch.no_grad():
            y1, f1 = encoder(img1)
            f1n, _ = recnet(y1)
            y2, f2 = encoder(img2)
            f2n, _ = recnet(y2)
        d_new.append(pair_cosine(f1n, f2n))
        d_raw.append(pair_cosine(f1, f2))
        lab += data["label"].tolist()
        idx += data["idx"].tolist()
    d_new = torch.cat(d_new).tolist()          # one device->host read for the whole set
    d_raw = torch.cat(d_raw).tolist()
    return np.array([d_new, lab, idx]).T, np.array([d_raw, lab, idx]).T


def get_fold_accuracy(fold, predicts, new=0):
    """lfw_eval.py:255-259 for ONE fold (kept for API parity; get_avg_accuracy sweeps all folds in one launch)."""
    n = predicts.shape[0]
    test = fold[1]
    n_folds = max(1, round(n / max(1, len(test))))
    bounds = [i * n // n_folds for i in range(n_folds + 1)]     # KFold's contiguous test ranges (lfw_eval.py:113-117)
    if test[0] not in bounds[:-1]:
        raise ValueError("fold does not start at a KFold boundary of %d folds over %d pairs" % (n_folds, n))
    f = bounds.index(test[0])
    res = _sweep_np(predicts, n_folds)
    return res["best_thr"][f], res["test_acc"][f]


def _sweep_np(predicts, n_folds):
    scores = torch.from_numpy(np.ascontiguousarray(predicts[:, 0]).astype(np.float32)).cuda()
    labels = torch.from_numpy(np.ascontiguousarray(predicts[:, 1]).astype(np.int32)).cuda()
    return threshold_sweep(scores, labels, n_folds)


def get_avg_accuracy(encoder, recnet, data_loader, flag=0, verbose=False):
    """lfw_eval.py:272-287 -> (avg_acc_new, avg_acc)."""
    pred_new, pred = calculate_distance(data_loader, encoder, recnet, flag)
    n_folds = 10
    r_new, r = _sweep_np(pred_new, n_folds), _sweep_np(pred, n_folds)
    if verbose:
        for t, a in zip(r["best_thr"], r["test_acc"]):
            print("Best threshold: {:.4f}; Test accuracy: {:.4f}".format(t, a))
        print("Average accuracy: {}".format(r["avg_acc"]))
    return r_new["avg_acc"], r["avg_acc"]


# ----------------------------------------------------------------------------------------------------------
# 1:N gallery scoring: the paired cosine of calculate_distance (lfw_eval.py:246-249) generalised
# to a full probe x gallery similarity matrix on the tcgen05 GEMM, rank-1 identification from its epilogue, and the
# accept counts for ROC / TAR@FAR with the reference's decision rule ((double)score > threshold on the np.arange grid).
# ----------------------------------------------------------------------------------------------------------
def _ceil(a, b):
    return (a + b - 1) // b * b


def gallery_cosine(probe, gallery, want_rank1=True):
    """probe (P,512), gallery (G,512) fp32 CUDA -> (cos (P,G) fp32 [a view of a padded matrix], rank1 (P,) int64 or None).
    Rows are L2-normalised as F.normalize does (for non-degenerate embeddings equal to f1.f2/(|f1||f2|+1e-8) of
    lfw_eval.py:246 to fp32 round-off); operands are split into bf16 hi + lo so the cosines carry ~1e-5 error."""
    if not (probe.is_cuda and gallery.is_cuda):
        raise RuntimeError("ffr_net_b200.scoring runs only on CUDA; there is no CPU fallback")
    probe, gallery = probe.contiguous().float(), gallery.contiguous().float()
    assert probe.dim() == 2 and gallery.dim() == 2 and probe.shape[1] == 512 and gallery.shape[1] == 512
    lib, P, dev = _lib.load(), _lib.ptr, probe.device
    n_p, n_g = probe.shape[0], gallery.shape[0]
    if n_p == 0 or n_g == 0:
        return torch.empty(n_p, n_g, dtype=torch.float32, device=dev), (torch.empty(n_p, dtype=torch.int64, device=dev)
                                                                         if want_rank1 else None)
    p_pad, g_pad = _ceil(n_p, 64), _ceil(n_g, 256)
    pp = torch.empty(p_pad, 1536, dtype=torch.bfloat16, device=dev)
    gp = torch.empty(g_pad, 1536, dtype=torch.bfloat16, device=dev)
    st = _lib.stream_ptr()
    _lib.check(lib.ffr_cosface_pack(P(probe), n_p, p_pad, 0, P(pp), None, 0, st), "pack(probe)")
    _lib.check(lib.ffr_cosface_pack(P(gallery), n_g, g_pad, 1, P(gp), None, 0, st), "pack(gallery)")
    cos = torch.empty(n_p, g_pad, dtype=torch.float32, device=dev)
    key = torch.empty(n_p, dtype=torch.int64, device=dev) if want_rank1 else None
    _lib.check(lib.ffr_gallery_cosine(P(pp), n_p, P(gp), g_pad, n_g, P(cos), P(key), st), "gallery_cosine")
    rank1 = None
    if want_rank1:
        rank1 = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    return cos[:, :n_g], rank1


def roc_counts(scores, probe_ids, gallery_ids, thresholds=None):
    """Accepted-pair counts at every threshold of the grid: returns dict(thresholds, true_accept, false_accept,
    n_genuine, n_impostor, tar, far) with integer counts (numpy int64) — TA(t) = #{genuine pairs with score > t}.
    scores: (P,G) fp32 CUDA (may be a strided view), ids: integer tensors."""
    if not scores.is_cuda:
        raise RuntimeError("ffr_net_b200.scoring runs only on CUDA; there is no CPU fallback")
    thr = thresholds_grid() if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    assert np.all(np.diff(thr) > 0), "thresholds must be ascending"
    dev = scores.device
    if scores.stride(1) != 1:
        scores = scores.contiguous()
    n_p, n_g = scores.shape
    pid = probe_ids.to(device=dev, dtype=torch.int32).contiguous()
    gid = gallery_ids.to(device=dev, dtype=torch.int32).contiguous()
    thr_d = torch.from_numpy(thr).to(dev)
    hist = torch.empty(2, len(thr) + 1, dtype=torch.int64, device=dev)
    lib = _lib.load()
    _lib.check(lib.ffr_roc_hist(_lib.ptr(scores), scores.stride(0), n_p, n_g, _lib.ptr(pid), _lib.ptr(gid),
                                _lib.ptr(thr_d), len(thr), _lib.ptr(hist), _lib.stream_ptr()), "roc_hist")
    h = hist.cpu().numpy()
    # bin b = number of thresholds below the score; accepted at threshold index t  <=>  b >= t + 1
    suffix = np.cumsum(h[:, ::-1], axis=1)[:, ::-1]
    ta, fa = suffix[0, 1:], suffix[1, 1:]
    n_gen, n_imp = int(h[0].sum()), int(h[1].sum())
    return dict(thresholds=thr, true_accept=ta, false_accept=fa, n_genuine=n_gen, n_impostor=n_imp,
                tar=ta / max(n_gen, 1), far=fa / max(n_imp, 1))


def tar_at_far(roc, far_target):
    """Largest TAR among the grid thresholds whose FAR does not exceed far_target (and that threshold)."""
    ok = np.nonzero(roc["far"] <= far_target)[0]
    if len(ok) == 0:
        return 0.0, None
    i = ok[np.argmax(roc["tar"][ok])]
    return float(roc["tar"][i]), float(roc["thresholds"][i])
