"""BASELINE.json configs[4] / SURVEY.md section 8c last row: LFW-style verification end to end on 600 synthetic pairs —
device pipeline (bf16 IR-SE50 + RecNet eval, pair cosine, 10-fold sweep; ffr_net_b200/lfw.py) against the fp32 CPU oracle
on the same images and the same (briefly fitted) RecNet weights, so that the rectified embeddings SPREAD (cosines 0.4..1.0)
instead of collapsing onto one vector as with untrained weights.

Tolerances. Raw backbone embeddings: max |cos - cos_oracle| <= 1e-3 (north star; measured 3e-4). Rectified embeddings:
<= 5e-2 max, <= 5e-3 median over 600 pairs (measured 3.3e-2 / 2.2e-3): the bf16 backbone's feature-map error (<= 1e-2 relative, the
north-star embedding tolerance) is amplified by RecNet itself — the fp32 ORACLE RecNet fed with the device backbone's maps
deviates by 2.6e-2 max / 1.9e-3 median from the all-fp32 result (tools/lfw_error_sources.py,
profiles/r02_lfw_error_sources.json), i.e. this is the conditioning of the rectifier, not a kernel defect; with
collapsed (untrained) embeddings the same pipeline agrees to 3e-5. Decisions at the oracle-chosen threshold of every fold
are identical for all pairs whose oracle score is farther from the threshold than the feature set's cosine tolerance."""
import numpy as np
import pytest
import torch

from oracle import backbone as ob
from oracle import recnet as orr
from oracle import scoring as osc
from ffr_net_b200 import lfw
from ffr_net_b200.recnet import RecNet
from ffr_net_b200.trainer import Trainer, default_opts

pytestmark = pytest.mark.gpu
N_PAIRS = 600


def test_lfw_600_pairs_decisions_match_oracle(lib):
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(lr=1e-3), recnet=rec, encoder_weights=bsd)
    lfw.fit_recnet(tr, steps=50, batch=32)                 # spread the rectified embeddings (deterministic)
    rec.eval()
    fitted = {k: v.detach().cpu().clone() for k, v in rec.state_dict().items()}
    img1, img2 = lfw.synth_pairs(0, N_PAIRS, 0, N_PAIRS)
    res = lfw.verify(tr.encoder, rec, n_pairs=N_PAIRS, images=(img1, img2), batch=200)
    torch.cuda.synchronize()
    labels = lfw.pair_labels(N_PAIRS).numpy()
    # ---- fp32 CPU oracle on the same images and weights ----
    s_new_ref, s_raw_ref = [], []
    with torch.no_grad():
        for i in range(0, N_PAIRS, 100):
            y1, f1 = ob.backbone_forward(bsd, img1[i:i + 100])
            y2, f2 = ob.backbone_forward(bsd, img2[i:i + 100])
            v1, _ = orr.recnet_forward(fitted, y1)
            v2, _ = orr.recnet_forward(fitted, y2)
            s_new_ref.append(osc.pair_cosine(v1, v2))
            s_raw_ref.append(osc.pair_cosine(f1, f2))
    s_new_ref, s_raw_ref = torch.cat(s_new_ref).numpy(), torch.cat(s_raw_ref).numpy()
    for name, got, ref in (("rectified", res["scores_rectified"].cpu().numpy(), s_new_ref),
                           ("raw", res["scores_raw"].cpu().numpy(), s_raw_ref)):
        err = np.abs(got - ref).max()
        sweep_ref = osc.sweep(ref, labels, 10)
        sweep_got = osc.sweep(got, labels, 10)
        print("%s: max |dcos| %.2e | oracle acc %.4f device acc %.4f | same: mean %.3f min %.3f, different: mean %.3f max %.3f"
              % (name, err, sweep_ref["avg_acc"], sweep_got["avg_acc"], ref[labels == 1].mean(), ref[labels == 1].min(),
                 ref[labels == 0].mean(), ref[labels == 0].max()))
        tol = 1e-3 if name == "raw" else 5e-2
        med = float(np.median(np.abs(got - ref)))
        print("%s: median |dcos| %.2e (tolerance: max %.0e)" % (name, med, tol))
        assert err <= tol and med <= tol / 10, name
        if name == "rectified":
            assert 0.55 < sweep_ref["avg_acc"] < 0.999, sweep_ref["avg_acc"]        # a discriminating, non-trivial task
            assert ref.max() - ref.min() > 0.1                                     # the embeddings spread
        # decisions at the oracle-chosen threshold of each fold, on that fold's held-out pairs
        per = N_PAIRS // 10
        flips = flips_all = ambiguous = 0
        for f, thr in enumerate(sweep_ref["best_thr"]):
            sl = slice(f * per, (f + 1) * per)
            d_ref, d_got = ref[sl].astype(np.float64) > thr, got[sl].astype(np.float64) > thr
            near = np.abs(ref[sl].astype(np.float64) - thr) <= tol
            ambiguous += int(near.sum())
            flips += int(((d_ref != d_got) & ~near).sum())
            flips_all += int((d_ref != d_got).sum())
        print("%s: %d pairs within %.0e of their threshold, %d decision flips outside that band, %d flips in all"
              % (name, ambiguous, tol, flips, flips_all))
        assert flips == 0                          # every disagreement is explained by the cosine tolerance
        # ... and the statement is not vacuous. The brief fit is chaotic: depending on rounding details of the training
        # kernels the fitted rectifier spreads the cosines over 0.4..1.0 (worst |dcos| 3e-2) or only over ~0.1 (worst
        # |dcos| 3e-3), so the fixed 5e-2 band can cover most pairs. Scale-aware form: at least half of the pairs are
        # decided with a margin of more than 1.5x the WORST cosine error of this run, and the decisions agree on
        # >= 97 % of ALL pairs.
        margin = 1.5 * float(err)
        decided = 0
        for f, thr in enumerate(sweep_ref["best_thr"]):
            sl = slice(f * per, (f + 1) * per)
            decided += int((np.abs(ref[sl].astype(np.float64) - thr) > margin).sum())
        print("%s: %d of %d pairs farther than 1.5 x max |dcos| = %.1e from their threshold" % (name, decided, N_PAIRS, margin))
        assert decided >= N_PAIRS // 2
        assert flips_all <= 0.03 * N_PAIRS
        assert abs(sweep_got["avg_acc"] - sweep_ref["avg_acc"]) <= 0.02
    # the device sweep itself (ffr_threshold_sweep) reproduces the oracle sweep on the device's own scores bit for bit
    for key, got in (("sweep_rectified", res["scores_rectified"]), ("sweep_raw", res["scores_raw"])):
        ref = osc.sweep(got.cpu().numpy(), labels, 10)
        assert res[key]["best_thr"] == ref["best_thr"] and res[key]["test_correct"] == ref["test_correct"]
