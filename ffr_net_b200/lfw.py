"""LFW-style verification on synthetic faces, whole box (BASELINE.json configs[4]; lfw/lfw_eval.py:226-287 as driven by
train.py:103-107): batched embedding (frozen IR-SE50 + RecNet), pair cosine for the rectified and the raw features,
10-fold x 400-threshold sweep. The 6000 pairs are sharded over the ranks (one process per GPU, no data-path collective);
the 2 x 6000 fp32 scores (48 KB) are gathered to rank 0, which runs both sweeps.

Pair structure mirrors pairs.txt (data/dataset.py:36-53): per fold of 600, 300 'same' pairs then 300 'different' ones.
A 'same' pair is an image and its masked copy (rows 56..111 overwritten by a per-image colour) with
a little noise; a 'different' pair is two independent images. Images are generated on the host from a seed per pair, so
every rank (and the CPU oracle in the tests) sees exactly the same data.
"""
import torch

from . import scoring

PAIRS = 6000
FOLD = 600


def pair_labels(n_pairs=PAIRS):
    per = max(2, n_pairs // 10)
    return torch.tensor([1 if (i % per) < per // 2 else 0 for i in range(n_pairs)], dtype=torch.int32)


_BLOCK = 50      # pairs are generated in aligned blocks of 50 from one seeded generator each


def identity_faces(g, n, pattern=None):
    """n synthetic 'faces' with identity-specific low-frequency structure: a random 3x7x7 field upsampled to 112x112
    (the identity) plus pixel noise (the photo). Independent Gaussian-noise images all have the same statistics and a
    network maps them to nearly the same embedding, so they cannot play different identities."""
    if pattern is None:
        pattern = torch.randn(n, 3, 7, 7, generator=g)
    up = torch.nn.functional.interpolate(pattern, size=(112, 112), mode="bilinear", align_corners=False)
    x = (0.8 * up + 0.2 * torch.randn(n, 3, 112, 112, generator=g)).clamp_(-1, 1)
    return x, pattern


def mask_rows(x, g, first_row=56):
    """The synthetic occlusion of SURVEY.md section 8d: rows first_row..111 overwritten by a per-image, per-channel colour."""
    y = x.clone()
    y[:, :, first_row:, :] = torch.empty(x.shape[0], 3, 1, 1).uniform_(-1, 1, generator=g)
    return y


def mask_lower_half(x, g):
    return mask_rows(x, g, 56)


def _synth_block(blk, seed, lab):
    """With random-init weights (no checkpoint is available offline) neither network carries identity information: what
    moves the cosine of a pair is how much of the second photo is occluded. The two classes are therefore built to be
    separated THROUGH the occlusion: 'same' = another photo of the same identity with a light occlusion (rows 84..111),
    'different' = a photo of another identity with a heavy one (rows 28..111, random start within 28..56). The scores
    spread over a wide range and the classes overlap a little, so the threshold sweep has an interior optimum."""
    g = torch.Generator().manual_seed(seed * 1000003 + blk)
    x, pat = identity_faces(g, _BLOCK)
    same_id, _ = identity_faces(g, _BLOCK, pat)
    same_id = mask_rows(same_id, g, 84)
    other, _ = identity_faces(g, _BLOCK)
    start = torch.randint(28, 57, (_BLOCK,), generator=g)
    col = torch.empty(_BLOCK, 3, 1, 1).uniform_(-1, 1, generator=g)
    rows = torch.arange(112).view(1, 1, 112, 1)
    other = torch.where(rows >= start.view(-1, 1, 1, 1), col, other)
    y = torch.where(lab.view(-1, 1, 1, 1).bool(), same_id, other)
    return x, y


def synth_pairs(lo, hi, seed=0, n_pairs=PAIRS):
    """Pairs [lo, hi) of the synthetic set: (img1, img2) fp32 (n,3,112,112) in [-1,1] (CPU tensors)."""
    lab = pair_labels(n_pairs)
    a = torch.empty(hi - lo, 3, 112, 112)
    b = torch.empty(hi - lo, 3, 112, 112)
    for blk in range(lo // _BLOCK, (hi + _BLOCK - 1) // _BLOCK):
        p0 = blk * _BLOCK
        lb = torch.zeros(_BLOCK, dtype=torch.int32)
        m = min(_BLOCK, n_pairs - p0)
        lb[:m] = lab[p0:p0 + m]
        x, y = _synth_block(blk, seed, lb)
        s, e = max(lo, p0), min(hi, p0 + _BLOCK)
        a[s - lo:e - lo], b[s - lo:e - lo] = x[s - p0:e - p0], y[s - p0:e - p0]
    return a, b


def fit_batch(batch, g):
    """One training batch of the same family: (unmasked photo, masked photo of the same identity, identity label)."""
    x, pat = identity_faces(g, batch)
    y, _ = identity_faces(g, batch, pat)
    return x, mask_lower_half(y, g), torch.randint(0, 10575, (batch,), generator=g)


def fit_recnet(trainer, steps=60, batch=32, seed=1234):
    """A short, deterministic fit of a randomly initialised RecNet on synthetic (photo, occluded photo) pairs so that the
    rectified embeddings spread like a trained model's (an untrained RecNet maps every face to almost the same vector and
    the verification task degenerates). Uses the Trainer's own step; returns the trainer."""
    g = torch.Generator().manual_seed(seed)
    dev = next(trainer.recnet.parameters()).device
    for it in range(steps):
        a, b, label = fit_batch(batch, g)
        trainer.step(a.to(dev), b.to(dev), label.to(dev))
    return trainer


def embed_pairs(encoder, recnet, img1, img2, batch=500):
    """(rectified cosine, raw cosine) of the pairs, fp32 CUDA (lfw_eval.py:241-249)."""
    dev = next(recnet.parameters()).device
    s_new, s_raw = [], []
    with torch.no_grad():
        for i in range(0, img1.shape[0], batch):
            a, b = img1[i:i + batch].to(dev, non_blocking=True), img2[i:i + batch].to(dev, non_blocking=True)
            n = a.shape[0]
            y, f = encoder(torch.cat((a, b)))                 # one backbone pass over both sides of the pairs
            v, _ = recnet(y)
            s_new.append(scoring.pair_cosine(v[:n].contiguous(), v[n:].contiguous()))
            s_raw.append(scoring.pair_cosine(f[:n].contiguous(), f[n:].contiguous()))
    return torch.cat(s_new), torch.cat(s_raw)


def verify(encoder, recnet, n_pairs=PAIRS, seed=0, batch=500, rank=0, world=1, images=None):
    """Whole pipeline for this rank's shard; rank 0 returns dict(acc_rectified, acc_raw, thresholds, scores...), the other
    ranks None. `images`: optional pre-generated (img1, img2) of this rank's shard (pinned host tensors)."""
    import torch.distributed as dist
    lo, hi = rank * n_pairs // world, (rank + 1) * n_pairs // world
    img1, img2 = images if images is not None else synth_pairs(lo, hi, seed, n_pairs)
    s_new, s_raw = embed_pairs(encoder, recnet, img1, img2, batch)
    mine = torch.stack((s_new, s_raw))                          # (2, shard)
    if world > 1:
        sizes = [(r + 1) * n_pairs // world - r * n_pairs // world for r in range(world)]
        pad = max(sizes)
        buf = torch.zeros(2, pad, dtype=torch.float32, device=mine.device)
        buf[:, : mine.shape[1]] = mine
        parts = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, parts, dst=0)
        if rank != 0:
            return None
        mine = torch.cat([p[:, :sz] for p, sz in zip(parts, sizes)], 1)
    labels = pair_labels(n_pairs).to(mine.device)
    r_new = scoring.threshold_sweep(mine[0].contiguous(), labels, 10)
    r_raw = scoring.threshold_sweep(mine[1].contiguous(), labels, 10)
    return {"acc_rectified": r_new["avg_acc"], "acc_raw": r_raw["avg_acc"], "sweep_rectified": r_new, "sweep_raw": r_raw,
            "scores_rectified": mine[0], "scores_raw": mine[1], "labels": labels}
