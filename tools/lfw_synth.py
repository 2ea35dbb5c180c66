"""BASELINE config[4]: LFW-style 6000-pair verification on synthetic faces — batched embedding (IR-SE50 + RecNet),
pair cosine and the 10-fold x 400-threshold sweep, all on the device. Prints one JSON object (also written to
gpurun_out/lfw_synth.json). Pair structure mirrors pairs.txt (data/dataset.py:36-53): per fold of 600, 300 'same'
(img2 = masked copy of img1 plus noise) then 300 'different' (independent image)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob   # synthetic weights / generators only
from ffr_net_b200 import synth as orr
from ffr_net_b200 import scoring
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet


def main():
    n_pairs, bs = 6000, 500
    dev = torch.device("cuda")
    enc = Backbone(50, 0.6, "ir_se")
    enc.load_state_dict(ob.synth_backbone_state_dict(0))
    enc = enc.to(dev).eval()
    rec = RecNet()
    rec.load_state_dict(orr.synth_recnet_state_dict(0))
    rec = rec.to(dev).eval()
    labels = torch.tensor(([1] * 300 + [0] * 300) * 10, dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)

    def batch(i):
        a = torch.randn(bs, 3, 112, 112, generator=g, device=dev).mul_(0.5).clamp_(-1, 1)
        other = torch.randn(bs, 3, 112, 112, generator=g, device=dev).mul_(0.5).clamp_(-1, 1)
        same = labels[i:i + bs].view(-1, 1, 1, 1).bool()
        b = torch.where(same, a + 0.1 * other, other).clamp_(-1, 1)
        b[:, :, 56:, :] = torch.where(same, torch.zeros_like(b[:, :, 56:, :]) + 0.3, b[:, :, 56:, :])   # "mask" on same pairs
        return a, b

    def run():
        s_new, s_raw = [], []
        with torch.no_grad():
            for i in range(0, n_pairs, bs):
                a, b = batch(i)
                y1, f1 = enc(a)
                v1, _ = rec(y1)
                v1 = v1.clone()
                f1 = f1.clone()
                y2, f2 = enc(b)
                v2, _ = rec(y2)
                s_new.append(scoring.pair_cosine(v1, v2))
                s_raw.append(scoring.pair_cosine(f1, f2))
        s_new, s_raw = torch.cat(s_new), torch.cat(s_raw)
        r_new = scoring.threshold_sweep(s_new, labels, 10)
        r_raw = scoring.threshold_sweep(s_raw, labels, 10)
        return s_new, r_new, r_raw

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s_new, r_new, r_raw = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # sweep alone
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        scoring.threshold_sweep(s_new, labels, 10)
    e1.record()
    torch.cuda.synchronize()
    out = {"pairs": n_pairs, "seconds_end_to_end": dt, "pairs_per_s": n_pairs / dt, "images_per_s": 2 * n_pairs / dt,
           "sweep_ms_incl_host_readback": e0.elapsed_time(e1) / 10, "acc_rectified": r_new["avg_acc"],
           "acc_raw": r_raw["avg_acc"], "best_thr_rectified": r_new["best_thr"][:3],
           "note": "includes on-device synthetic image generation; reference sweep alone is 15 s of Python loops (SURVEY §6)"}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/lfw_synth.json", "w"), indent=1)


if __name__ == "__main__":
    main()
