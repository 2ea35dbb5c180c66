"""Generates tests/golden/*.npz by running the REAL reference (imported from /root/reference, read-only) on the
deterministic synthetic weights/inputs of oracle/. Run in the build container only:

    python tools/make_golden.py

The fixtures pin the oracle (and, through it, the CUDA path) to the reference's behaviour; the reference itself ships
no tests or golden vectors. Modules the reference imports but this image lacks (imageio, skimage,
tensorboardX, matplotlib) are stubbed — they are not on the hot path.
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import backbone as ob  # noqa: E402
from oracle import recnet as orr  # noqa: E402
from oracle import scoring as osc  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Stub(types.ModuleType):
    """Module stand-in: any attribute is a no-op callable/class (the stubbed modules are never used on the hot path)."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})


def stub_missing():
    for name in ("imageio", "skimage", "skimage.transform", "skimage.io", "tensorboardX", "matplotlib",
                 "matplotlib.pyplot", "h5py"):
        ok = False
        if name not in ("imageio",):
            try:
                __import__(name)
                ok = True
            except Exception:
                ok = False
        if not ok:
            sys.modules[name] = _Stub(name)


def golden_preprocess():
    """data/dataset.py CASIA.__getitem__ (the REAL class, on a temporary PNG dataset) + the transform of
    data/dataloader.py:15-19 -> tests/golden/preprocess_ref.npz."""
    import random
    import shutil
    import tempfile
    from PIL import Image
    from torchvision import transforms
    from oracle import preprocess as opp
    try:
        import cv2  # noqa: F401
    except Exception:
        sys.modules["cv2"] = _Stub("cv2")              # imported at module top by data/dataset.py, unused on this path
    from data import dataset as rds
    tmp = tempfile.mkdtemp(prefix="ffr_golden_ds_")
    try:
        size = 16
        imgs = opp.synth_images_u8(4, size, seed=5)
        masks = opp.synth_images_u8(4, size, seed=6)
        os.makedirs(os.path.join(tmp, "id0"))
        lines = []
        for i in range(4):
            os.makedirs(os.path.join(tmp, "p%d" % i))
            Image.fromarray(imgs[i]).save(os.path.join(tmp, "p%d" % i, "%03d.png" % i))
            Image.fromarray(masks[i]).save(os.path.join(tmp, "p%d" % i, "%03d_mask.png" % i))
            lines.append("p%d/%03d.png %d" % (i, i, 100 + i))
        lst = os.path.join(tmp, "list.txt")
        open(lst, "w").write("\n".join(lines) + "\n")
        tf_transform = transforms.Compose([transforms.ToTensor(),                      # data/dataloader.py:15-19
                                           transforms.Normalize([0.5, 0.5, 0.5], [0.5, 0.5, 0.5])])
        ds = rds.CASIA(tmp, lst, transform=tf_transform, shuffle=False, size=(size, size))
        out1, out2, flips, labels = [], [], [], []
        for i in range(4):
            random.seed(40 + i)
            flips.append(random.random() < 0.5)        # the draw __getitem__ makes (dataset.py:147-149)
            random.seed(40 + i)
            sample = ds[i]
            out1.append(sample["img1"].numpy())
            out2.append(sample["img2"].numpy())
            labels.append(int(sample["label"]))
        assert any(flips) and not all(flips), flips
        np.savez_compressed(os.path.join(OUT, "preprocess_ref.npz"), imgs=imgs, masks=masks,
                            flips=np.array(flips, dtype=np.uint8), img1=np.stack(out1), img2=np.stack(out2),
                            labels=np.array(labels))
        print("preprocess_ref.npz: flips", flips, "labels", labels)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


GOLD_TRAIN_TENSORS = ["classifier.weight", "Conv4Merge.0.conv2d.weight", "Conv4Merge.1.conv2.norm.norm.bias",
                      "ChannelFlipMerge.0.conv2d.weight", "Conv4Space.0.conv2d.weight", "Conv4Space.0.norm.norm.weight",
                      "Conv4Space.5.conv2.relu.func.weight", "Conv4Channel.0.weight", "Conv4Channel.8.bias",
                      "Conv4Channel.4.func.weight"]


def gold_slice(t):
    """Deterministic thinning that keeps fixtures small: every 997th element of the flattened tensor (all if small)."""
    f = t.detach().reshape(-1)
    return f.clone() if f.numel() <= 2048 else f[::997].clone()


def golden_trainer():
    """The REAL models/trainer.py Trainer (set_input / forward / backward / optimizer_parameters, :129-187) on a
    hand-built instance (its __init__ wants the Google-Drive checkpoint and GPU ids), CPU fp32, 2 pairs:
    losses, accuracy, gradient slices before clipping, parameters after clip + Adam, BN running statistics
    -> tests/golden/trainer_ref.npz. recnet.py:262 hard-codes device='cuda' for the one-hot; torch.zeros is patched to
    drop that keyword (no other change)."""
    stub_missing()
    try:
        import cv2  # noqa: F401
    except Exception:
        sys.modules["cv2"] = _Stub("cv2")
    import types as _types
    from pretrain.model_ir_se50 import Backbone
    import models.recnet as mr
    import models.trainer as mt
    from ffr_net_b200 import synth
    real_zeros = torch.zeros

    def zeros_cpu(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    bsd, rsd = synth.synth_backbone_state_dict(0), synth.synth_recnet_state_dict(0)
    img1, img2 = synth.synth_faces(2, seed=3), synth.synth_faces(2, seed=3, masked=True)
    label = torch.tensor([5, 4242])
    tr = mt.Trainer.__new__(mt.Trainer)
    tr.opts = _types.SimpleNamespace(phase="train", lr=1e-3, beta1=0.9, beta2=0.999, weight_decay=0.0, optimizer="adam",
                                     loss_weight=[1.0, 1.0, 1.0, 1.0])
    tr.isTrain, tr.lr = True, tr.opts.lr
    tr.encoder = Backbone(50, 0.6, "ir_se")
    tr.encoder.load_state_dict(bsd, strict=True)
    tr.recnet = mr.RecNet(norm_type="bn", relu_type="prelu")
    tr.recnet.load_state_dict(rsd, strict=True)
    for p in tr.encoder.parameters():
        p.requires_grad = False
    tr.forward_encoder = lambda x: tr.encoder(x)
    tr.forward_recnet = lambda x, l: tr.recnet(x, l)
    tr.encoder.eval()
    tr.recnet.train()
    tr.config_optimizer()
    tr.config_criterion()
    mr.torch.zeros = zeros_cpu
    try:
        tr.set_input(img1, img2, label)
        with torch.no_grad():
            pass
        tr.forward()
        tr.optim.zero_grad()
        tr.backward()
        named = dict(tr.recnet.named_parameters())
        grads = {k: gold_slice(named[k].grad) for k in GOLD_TRAIN_TENSORS}
        gnorm = {k: float(named[k].grad.norm()) for k in named}
        losses = [float(l.detach()) for l in tr.loss_items]
        # the optimizer half of optimizer_parameters (:183-187) on the gradients just computed
        mt.clip_grad_value_(tr.recnet.parameters(), 1.0)
        tr.optim.step()
        after = {k: gold_slice(named[k]) for k in GOLD_TRAIN_TENSORS}
    finally:
        mr.torch.zeros = real_zeros
    sd = tr.recnet.state_dict()
    out = {"losses": np.array(losses), "accuracy": np.float64(tr.accuracy), "pos_loss": np.float64(float(tr.pos_loss)),
           "neg_loss": np.float64(float(tr.neg_loss)),
           "grad_norm_keys": np.array(sorted(gnorm)), "grad_norms": np.array([gnorm[k] for k in sorted(gnorm)]),
           "run_mean_merge0": sd["Conv4Merge.0.norm.norm.running_mean"].numpy(),
           "run_var_space0": sd["Conv4Space.0.norm.norm.running_var"].numpy(),
           "nbt": np.int64(int(sd["Conv4Merge.0.norm.norm.num_batches_tracked"]))}
    for k in GOLD_TRAIN_TENSORS:
        out["grad:" + k] = grads[k].numpy()
        out["after:" + k] = after[k].numpy()
    np.savez_compressed(os.path.join(OUT, "trainer_ref.npz"), **out)
    print("trainer_ref.npz: losses", losses, "acc", tr.accuracy)


def golden_checkpoint():
    """A tiny checkpoint written by the REAL utils.save (utils/utils.py:110-115) with the container layout of
    Trainer.save_model (models/trainer.py:216-224) -> tests/golden/tiny_ckpt_ref.pth.gzip."""
    stub_missing()
    try:
        import cv2  # noqa: F401
    except Exception:
        sys.modules["cv2"] = _Stub("cv2")
    from utils import utils as ru
    g = torch.Generator().manual_seed(11)
    obj = {"RecNet": {"a.weight": torch.randn(3, 4, generator=g), "a.norm.num_batches_tracked": torch.tensor(7)},
           "optimizer": {"state": {}, "param_groups": [{"lr": 0.05, "betas": (0.9, 0.999), "params": [0]}]},
           "epoch": 3, "iter": 1234}
    ru.save(obj, os.path.join(OUT, "tiny_ckpt_ref.pth.gzip"))


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        return golden_preprocess()
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        return golden_checkpoint()
    if len(sys.argv) > 1 and sys.argv[1] == "trainer":
        return golden_trainer()
    torch.set_num_threads(os.cpu_count() or 1)
    from pretrain.model_ir_se50 import Backbone
    import models.recnet as mr

    # ---- backbone ----
    sd = ob.synth_backbone_state_dict(0)
    ref = Backbone(50, 0.6, "ir_se").eval()
    ref.load_state_dict(sd, strict=True)
    x = ob.synth_faces(2, seed=1)
    xm = ob.synth_faces(2, seed=1, masked=True)
    with torch.no_grad():
        y, f = ref(x)
        ym, fm = ref(xm)
    np.savez_compressed(os.path.join(OUT, "backbone_ref.npz"), f=f.numpy(), f_masked=fm.numpy(),
                        y_slice=y[:, ::64, :, :].numpy(), y_abs_sum=np.float64(y.double().abs().sum().item()),
                        ym_abs_sum=np.float64(ym.double().abs().sum().item()))

    # ---- RecNet eval + train-mode (label) forward ----
    rsd = orr.synth_recnet_state_dict(0)
    rec = mr.RecNet()
    rec.load_state_dict(rsd, strict=True)
    g = torch.Generator().manual_seed(3)
    fx = torch.randn(3, 512, 7, 7, generator=g) * 0.3
    label = torch.randint(0, 10575, (3,), generator=g)
    rec.eval()
    with torch.no_grad():
        v, fmap = rec(fx)
    _z = torch.zeros
    mr.torch.zeros = lambda *a, **k: _z(*a, **{kk: vv for kk, vv in k.items() if kk != "device"})  # recnet.py:262
    rec.train()
    with torch.no_grad():
        out = rec(fx, label)
    mr.torch.zeros = _z
    tsd = rec.state_dict()
    np.savez_compressed(
        os.path.join(OUT, "recnet_ref.npz"), v=v.numpy(), fmap_slice=fmap[:, ::64].numpy(),
        label=label.numpy(), t_v=out[0].numpy(), t_pred_loss_slice=out[1][:, ::97].numpy(),
        t_pred_label_slice=out[2][:, ::97].numpy(), t_m_space=out[3].numpy(), t_m_channel_slice=out[4][:, ::37, ::41].numpy(),
        t_feat_space_slice=out[5][:, ::64].numpy(), t_feat_channel_slice=out[6][:, ::64].numpy(),
        t_run_mean_merge0=tsd["Conv4Merge.0.norm.norm.running_mean"].numpy(),
        t_run_var_merge0=tsd["Conv4Merge.0.norm.norm.running_var"].numpy(),
        t_nbt=np.int64(tsd["Conv4Merge.0.norm.norm.num_batches_tracked"].item()))

    # ---- selfSimilarity ----
    with torch.no_grad():
        ss_s, ss_c = mr.selfSimilarity(fx)
    np.savez_compressed(os.path.join(OUT, "selfsim_ref.npz"), ss_space=ss_s.numpy(), ss_channel_slice=ss_c[:, ::16, ::16].numpy())

    # ---- preprocessing, checkpoint container ----
    golden_preprocess()
    golden_checkpoint()
    golden_trainer()

    # ---- scoring: the real lfw_eval functions on 6000 synthetic pairs ----
    stub_missing()
    import lfw.lfw_eval as le
    scores, labels = osc.synth_pair_scores(6000, 0)
    pred = np.array([scores.astype(np.float64).tolist(), labels.tolist(), list(range(6000))]).T
    folds = le.KFold(n=6000, n_folds=10, shuffle=False)
    best, acc = [], []
    for fold in folds:
        b, a = le.get_fold_accuracy(fold, pred, 0)
        best.append(b)
        acc.append(a)
    f1 = torch.randn(16, 512, generator=g)
    f2 = torch.randn(16, 512, generator=g) + 0.5 * f1
    cos = torch.sum(f1 * f2, dim=1) / (f1.norm(dim=1) * f2.norm(dim=1) + 1e-8)     # lfw_eval.py:246 verbatim formula
    np.savez_compressed(os.path.join(OUT, "scoring_ref.npz"), scores=scores, labels=labels, best_thr=np.array(best),
                        test_acc=np.array(acc), avg_acc=np.float64(sum(acc) / 10), f1=f1.numpy(), f2=f2.numpy(),
                        cos=cos.numpy())
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
