"""CPU tests: the oracle restatement reproduces the golden vectors generated from the REAL reference
(tools/make_golden.py, run in the build container where /root/reference is mounted). Nothing here reads
/root/reference."""
import os

import numpy as np
import torch

from oracle import backbone as ob
from oracle import recnet as orr
from oracle import scoring as osc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_backbone_oracle_matches_reference_golden():
    g = _load("backbone_ref.npz")
    sd = ob.synth_backbone_state_dict(0)
    assert len(sd) == 402
    with torch.no_grad():
        y, f = ob.backbone_forward(sd, ob.synth_faces(2, seed=1))
        ym, fm = ob.backbone_forward(sd, ob.synth_faces(2, seed=1, masked=True))
    # same ops in the same order as the reference -> agreement to fp32 round-off
    assert np.abs(f.numpy() - g["f"]).max() <= 1e-6
    assert np.abs(fm.numpy() - g["f_masked"]).max() <= 1e-6
    assert np.abs(y[:, ::64].numpy() - g["y_slice"]).max() <= 1e-5
    assert abs(y.double().abs().sum().item() - float(g["y_abs_sum"])) <= 1e-6 * float(g["y_abs_sum"])
    assert abs(ym.double().abs().sum().item() - float(g["ym_abs_sum"])) <= 1e-6 * float(g["ym_abs_sum"])


def test_recnet_oracle_eval_matches_reference_golden():
    g = _load("recnet_ref.npz")
    sd = orr.synth_recnet_state_dict(0)
    assert len(sd) == 121
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(3, 512, 7, 7, generator=gen) * 0.3
    with torch.no_grad():
        v, fmap = orr.recnet_forward(sd, x)
    assert np.abs(v.numpy() - g["v"]).max() <= 1e-5 * np.abs(g["v"]).max()
    assert np.abs(fmap[:, ::64].numpy() - g["fmap_slice"]).max() <= 1e-5 * np.abs(g["fmap_slice"]).max()


def test_recnet_oracle_train_matches_reference_golden():
    g = _load("recnet_ref.npz")
    sd = orr.synth_recnet_state_dict(0)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(3, 512, 7, 7, generator=gen) * 0.3
    label = torch.randint(0, 10575, (3,), generator=gen)
    assert np.array_equal(label.numpy(), g["label"])
    with torch.no_grad():
        out, stats = orr.recnet_forward(sd, x, label, training=True, return_stats=True)

    def close(a, b, tol=2e-5):
        return np.abs(a - b).max() <= tol * max(1e-6, np.abs(b).max())
    assert close(out[0].numpy(), g["t_v"])
    assert close(out[1][:, ::97].numpy(), g["t_pred_loss_slice"])
    assert close(out[2][:, ::97].numpy(), g["t_pred_label_slice"])
    assert close(out[3].numpy(), g["t_m_space"])
    assert close(out[4][:, ::37, ::41].numpy(), g["t_m_channel_slice"])
    assert close(out[5][:, ::64].numpy(), g["t_feat_space_slice"])
    assert close(out[6][:, ::64].numpy(), g["t_feat_channel_slice"])
    assert close(stats["Conv4Merge.0.norm.norm.running_mean"].numpy(), g["t_run_mean_merge0"])
    assert close(stats["Conv4Merge.0.norm.norm.running_var"].numpy(), g["t_run_var_merge0"])
    assert int(stats["Conv4Merge.0.norm.norm.num_batches_tracked"]) == int(g["t_nbt"]) == 1


def test_selfsim_oracle_matches_reference_golden():
    g = _load("selfsim_ref.npz")
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(3, 512, 7, 7, generator=gen) * 0.3
    s, c = orr.self_similarity(x)
    assert s.shape == (3, 49, 7, 7) and c.shape == (3, 512, 512)
    assert np.abs(s.numpy() - g["ss_space"]).max() <= 1e-6
    assert np.abs(c[:, ::16, ::16].numpy() - g["ss_channel_slice"]).max() <= 1e-6


def test_scoring_oracle_matches_reference_golden():
    """The vectorised sweep reproduces the real lfw_eval.get_fold_accuracy on 6000 pairs: thresholds bit-exact."""
    g = _load("scoring_ref.npz")
    scores, labels = osc.synth_pair_scores(6000, 0)
    assert np.array_equal(scores, g["scores"]) and np.array_equal(labels, g["labels"])
    res = osc.sweep(scores, labels, 10)
    assert res["best_thr"] == g["best_thr"].tolist()
    assert res["test_acc"] == g["test_acc"].tolist()
    assert res["avg_acc"] == float(g["avg_acc"])
    cos = osc.pair_cosine(torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"]))
    assert np.abs(cos.numpy() - g["cos"]).max() <= 1e-7


def test_scoring_literal_equals_vectorised():
    """The pure-Python restatement (reference control flow) and the numpy one agree exactly, incl. ties."""
    for seed, n in ((1, 600), (2, 300)):
        s, l = osc.synth_pair_scores(600, seed)
        s, l = s[:n], l[:n]
        pred = np.array([s.astype(np.float64), l, np.arange(n)]).T
        lit = [osc.fold_accuracy_literal(f, pred) for f in osc.kfold(n, 10)]
        vec = osc.sweep(s, l, 10)
        assert [a for a, _ in lit] == vec["best_thr"]
        assert [b for _, b in lit] == vec["test_acc"]
    thr = osc.thresholds_grid()
    assert len(thr) == 400 and thr[0] == -1.0 and thr[259] != 0.295      # not a round decimal


def test_preprocess_oracle_matches_reference_golden():
    """oracle.preprocess reproduces the real data/dataset.py CASIA.__getitem__ + dataloader transform bit for bit
    (channel swap, flips drawn by the reference, ToTensor, Normalize)."""
    from oracle import preprocess as opp
    g = _load("preprocess_ref.npz")
    assert np.array_equal(g["imgs"], opp.synth_images_u8(4, 16, seed=5))
    got1 = opp.preprocess_batch(g["imgs"], g["flips"]).numpy()
    got2 = opp.preprocess_batch(g["masks"], g["flips"]).numpy()
    assert np.array_equal(got1, g["img1"]) and np.array_equal(got2, g["img2"])
    assert g["flips"].any() and not g["flips"].all()


def test_gallery_oracle_consistent_with_paired_scoring():
    """The 1:N oracle restates the paired rule: its diagonal equals oracle pair_cosine, and its counts at a threshold
    equal a literal Python loop of eval_acc's comparison (lfw_eval.py:141-147)."""
    g = torch.Generator().manual_seed(2)
    a, b = torch.randn(9, 512, generator=g), torch.randn(9, 512, generator=g)
    m = osc.gallery_cosine(a.numpy(), b.numpy())
    assert np.abs(np.diag(m) - osc.pair_cosine(a, b).numpy().astype(np.float64)).max() <= 1e-6
    pid, gid = np.arange(9) % 3, np.arange(9) % 3
    roc = osc.roc_counts(m.astype(np.float32), pid, gid, thresholds=[-0.05, 0.0, 0.05])
    for k, t in enumerate([-0.05, 0.0, 0.05]):
        ta = fa = 0
        for i in range(9):
            for j in range(9):
                same = 1 if float(np.float32(m[i, j])) > t else 0
                if same and pid[i] == gid[j]:
                    ta += 1
                if same and pid[i] != gid[j]:
                    fa += 1
        assert roc["true_accept"][k] == ta and roc["false_accept"][k] == fa


def test_train_oracle_matches_reference_trainer_golden():
    """oracle.train.train_step + clip_adam_step against the REAL models/trainer.py Trainer (set_input / forward /
    backward / clip_grad_value_ / Adam.step on a hand-built instance, tools/make_golden.py golden_trainer): the four
    weighted loss items, accuracy, the norm of every one of the 76 gradients, gradient and post-step parameter slices,
    BatchNorm running statistics after the two forward calls."""
    from oracle import train as otr
    g = _load("trainer_ref.npz")
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    img1, img2 = ob.synth_faces(2, seed=3), ob.synth_faces(2, seed=3, masked=True)
    label = torch.tensor([5, 4242])
    items, grads, stats, acc = otr.train_step(bsd, rsd, img1, img2, label)
    assert np.allclose(items, g["losses"], rtol=1e-5, atol=1e-7), (items, g["losses"])     # measured 9e-8
    assert acc == float(g["accuracy"])
    keys = [str(k) for k in g["grad_norm_keys"]]
    assert sorted(grads) == keys and len(keys) == 76
    norms = np.array([float(grads[k].norm()) for k in keys])
    assert np.allclose(norms, g["grad_norms"], rtol=1e-4, atol=1e-9), np.abs(norms / g["grad_norms"] - 1).max()   # 8e-7

    def sl(t):
        f = t.detach().reshape(-1)
        return f if f.numel() <= 2048 else f[::997]
    params = {k: rsd[k] for k in grads}
    after, _ = otr.clip_adam_step(params, grads, lr=1e-3)
    for name in [k[5:] for k in g.files if k.startswith("grad:")]:
        ref = g["grad:" + name]
        got = sl(grads[name]).numpy()
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-9, name
        ref_p, got_p = g["after:" + name], sl(after[name]).numpy()
        # Adam's first step moves every element by ~lr * sign(g): compare the step, not just the value
        ref_0 = sl(rsd[name]).numpy()
        flip = np.abs((got_p - ref_0) - (ref_p - ref_0)) > 2e-4 * 1e-3
        assert flip.mean() <= 0.02, (name, flip.mean())     # only elements whose gradient is ~0 may disagree
    assert np.allclose(stats["Conv4Merge.0.norm.norm.running_mean"].numpy(), g["run_mean_merge0"], rtol=1e-4, atol=1e-6)
    assert np.allclose(stats["Conv4Space.0.norm.norm.running_var"].numpy(), g["run_var_space0"], rtol=1e-4, atol=1e-6)
    assert int(stats["Conv4Merge.0.norm.norm.num_batches_tracked"]) == int(g["nbt"]) == 2
