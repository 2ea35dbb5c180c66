"""Oracle (test infrastructure): restatement of the LFW verification scoring of /root/reference/lfw/lfw_eval.py.

Two forms: a literal one (pure-Python loops, exactly the reference's control flow — small inputs only) and a
vectorised numpy float64 one (same decisions, used at full size). tests/ shows the two equal, and
tools/make_golden.py pins both against the real reference functions imported in the build container.
"""
import numpy as np


def thresholds_grid():
    """get_fold_accuracy, lfw_eval.py:256: np.arange(-1.0, 1.0, 0.005) — 400 float64 values that are NOT round
    decimals (e.g. 0.29500000000000126); decisions must use exactly these values."""
    return np.arange(-1.0, 1.0, 0.005)


def kfold(n=6000, n_folds=10):
    """KFold(shuffle=False), lfw_eval.py:110-118: contiguous test folds, train = the rest."""
    folds = []
    base = list(range(n))
    for i in range(n_folds):
        test = base[i * n // n_folds:(i + 1) * n // n_folds]
        train = sorted(set(base) - set(test))
        folds.append([train, test])
    return folds


def pair_cosine(f1, f2):
    """calculate_distance, lfw_eval.py:246,248: sum(f1*f2,1) / (|f1|*|f2| + 1e-8) in fp32 (torch tensors)."""
    import torch
    return torch.sum(f1 * f2, dim=1) / (f1.norm(dim=1) * f2.norm(dim=1) + 1e-8)


# ---- literal restatement (reference control flow) ----------------------------------------------------------
def eval_acc_literal(threshold, diff):
    """eval_acc, lfw_eval.py:137-153 (save_wrong=0): predicted same iff float(score) > threshold."""
    y_true, y_pred = [], []
    for d in diff:
        y_pred.append(1 if float(d[0]) > threshold else 0)
        y_true.append(int(d[1]))
    y_true, y_pred = np.array(y_true), np.array(y_pred)
    return 1.0 * np.count_nonzero(y_true == y_pred) / len(y_true)


def find_best_threshold_literal(thresholds, predicts):
    """find_best_threshold, lfw_eval.py:155-162: `>=` keeps the LAST threshold with the best accuracy."""
    best_threshold = best_acc = 0
    for threshold in thresholds:
        accuracy = eval_acc_literal(threshold, predicts)
        if accuracy >= best_acc:
            best_acc = accuracy
            best_threshold = threshold
    return best_threshold


def fold_accuracy_literal(fold, predicts):
    """get_fold_accuracy, lfw_eval.py:255-259."""
    thresholds = thresholds_grid()
    best = find_best_threshold_literal(thresholds, predicts[fold[0]])
    return best, eval_acc_literal(best, predicts[fold[1]])


# ---- vectorised restatement ---------------------------------------------------------------------------------
def sweep(scores, labels, n_folds=10, thresholds=None):
    """All folds at once. scores: fp32/fp64 array (n,), labels (n,) in {0,1}.
    Returns dict(best_idx, best_thr, test_correct, train_correct, test_acc, avg_acc) following
    lfw_eval.py:255-268,274-287 (avg divides by the literal 10 when n_folds == 10)."""
    thr = thresholds_grid() if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    s = np.asarray(scores).astype(np.float64)
    y = np.asarray(labels).astype(np.int64)
    n = len(s)
    pred = (s[None, :] > thr[:, None]).astype(np.int64)            # [T, n]
    correct = (pred == y[None, :]).astype(np.int64)
    bounds = [i * n // n_folds for i in range(n_folds + 1)]
    per_fold = np.stack([correct[:, bounds[f]:bounds[f + 1]].sum(1) for f in range(n_folds)], 0)   # [folds, T]
    total = per_fold.sum(0)
    out = dict(best_idx=[], best_thr=[], test_correct=[], train_correct=[], test_acc=[])
    for f in range(n_folds):
        train = total - per_fold[f]
        best = int(np.flatnonzero(train == train.max())[-1])        # last maximiser (>= rule)
        out["best_idx"].append(best)
        out["best_thr"].append(float(thr[best]))
        out["train_correct"].append(int(train[best]))
        out["test_correct"].append(int(per_fold[f, best]))
        out["test_acc"].append(1.0 * per_fold[f, best] / (bounds[f + 1] - bounds[f]))
    out["avg_acc"] = sum(out["test_acc"]) / (10 if n_folds == 10 else n_folds)
    return out


def synth_pair_scores(n=6000, seed=0):
    """Synthetic LFW-like scores/labels: per 600-fold the first 300 pairs 'same' (label 1), next 300 'different'
    (data/dataset.py:36-53 pairs.txt structure), scores drawn so the classes overlap a little."""
    rng = np.random.RandomState(seed)
    labels = np.zeros(n, dtype=np.int64)
    per = n // 10
    for f in range(10):
        labels[f * per:f * per + per // 2] = 1
    scores = np.where(labels == 1, rng.normal(0.55, 0.18, n), rng.normal(0.05, 0.15, n)).clip(-0.999, 0.999)
    return scores.astype(np.float32), labels


# ----------------------------------------------------------------------------------------------------------
# 1:N gallery scoring: the paired rule of lfw_eval.py:246-249 / eval_acc :141-147 applied to every (probe, gallery)
# pair. float64 cosines, literal counting loops over the threshold grid.
# ----------------------------------------------------------------------------------------------------------
def gallery_cosine(probe, gallery):
    """(P,D),(G,D) -> (P,G) float64: f1.f2 / (|f1||f2| + 1e-8), the formula of lfw_eval.py:246."""
    p = np.asarray(probe, dtype=np.float64)
    g = np.asarray(gallery, dtype=np.float64)
    num = p @ g.T
    den = np.linalg.norm(p, axis=1)[:, None] * np.linalg.norm(g, axis=1)[None, :] + 1e-8
    return num / den


def roc_counts(scores_f32, probe_ids, gallery_ids, thresholds=None):
    """Accepted genuine / impostor pair counts per threshold: predict same iff float(score) > threshold (strict),
    as eval_acc does (lfw_eval.py:141-147)."""
    thr = np.arange(-1.0, 1.0, 0.005) if thresholds is None else np.asarray(thresholds, dtype=np.float64)
    s = np.asarray(scores_f32, dtype=np.float32).astype(np.float64)
    same = np.asarray(probe_ids)[:, None] == np.asarray(gallery_ids)[None, :]
    ta = np.array([int(np.count_nonzero((s > t) & same)) for t in thr], dtype=np.int64)
    fa = np.array([int(np.count_nonzero((s > t) & ~same)) for t in thr], dtype=np.int64)
    return dict(thresholds=thr, true_accept=ta, false_accept=fa, n_genuine=int(same.sum()), n_impostor=int((~same).sum()))
