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
    f = test[0] * n_folds // n
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
