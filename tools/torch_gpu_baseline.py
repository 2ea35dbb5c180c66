"""Developer timing script (GPU box): the reference ALGORITHM through stock PyTorch GPU library ops (cuDNN / cuBLAS),
i.e. what the reference does on a GPU as shipped, next to the hand-written path. Uses the oracle's restatement of the
network (same ops the reference modules call) on the CUDA device: fp32 without TF32 (the true-fp32 arm), fp32 with
TF32 (PyTorch's GPU default for convolutions), and bf16 autocast + channels_last (strongest library baseline).
Not the contract bench: SURVEY.md 8(d) "CPU baseline beside it", last sentence."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone as ob
from oracle import recnet as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    bsd = ob.synth_backbone_state_dict(0)
    rsd = orr.synth_recnet_state_dict(0)
    x = ob.synth_faces(min(n, 64), 0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()
    bsd_g = {k: v.cuda() for k, v in bsd.items()}
    rsd_g = {k: v.cuda() for k, v in rsd.items()}

    def lib_path():
        y, _ = ob.backbone_forward(bsd_g, x)
        v, _ = orr.recnet_forward(rsd_g, y)
        return v

    out = {"n": n, "runs": []}
    with torch.no_grad():
        for name, tf32, amp in (("fp32 (TF32 off)", False, False), ("fp32 + TF32 (PyTorch GPU default for convs)", True, False),
                                ("bf16 autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            if amp:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ms = timed(lib_path)
            else:
                ms = timed(lib_path)
            row = {"arm": "stock PyTorch library ops, " + name, "ms": ms, "img_s": n / ms * 1e3}
            print(json.dumps(row))
            out["runs"].append(row)
        enc = Backbone(50, 0.6, "ir_se")
        enc.load_state_dict(bsd)
        enc = enc.cuda().eval()
        rec = RecNet()
        rec.load_state_dict(rsd)
        rec = rec.cuda().eval()
        ms = timed(lambda: rec.embed_from_images(enc, x), iters=10, warm=3)
        row = {"arm": "ffr_net_b200 (hand-written sm_100a path)", "ms": ms, "img_s": n / ms * 1e3}
        print(json.dumps(row))
        out["runs"].append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/torch_gpu_baseline_%d.json" % n, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
