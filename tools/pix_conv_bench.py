"""Developer timing script (GPU box): the RecNet 3x3 convolutions (eval epilogue: bias + PReLU + H9 scatter) at n images,
row-major tiles (sliding-window kernel over all 81 H9 rows) vs pixel-major tiles (128 images per pixel, 49 pixels).
Usage: python tools/pix_conv_bench.py [n] [only_cin only_cout]   (the optional pair restricts to one shape, for ncu)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib, packing
from ffr_net_b200.recnet import _h9_scatter

SHAPES = [(576, 256), (256, 256), (256, 128), (128, 128), (1024, 512), (512, 512), (1536, 512)]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    only = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else None
    lib = _lib.load()
    P = _lib.ptr
    tab = _h9_scatter(0, "cuda")
    rows = []
    for cin, cout in SHAPES:
        if only and (cin, cout) != only:
            continue
        g = torch.Generator(device="cuda").manual_seed(0)
        x = (torch.randn(n * 81, cin, generator=g, device="cuda") * 0.5).bfloat16()
        w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
        wp = packing.pack_conv(w)
        bias = torch.zeros(cout, device="cuda")
        slope = torch.full((cout,), 0.25, device="cuda")
        out = torch.zeros(n * 81, cout, dtype=torch.bfloat16, device="cuda")
        row = {"cin": cin, "cout": cout, "n": n}
        for mode, name in ((0, "rowmajor"), (1, "pixmajor")):
            lib.ffr_debug_set_pixmajor(mode)

            def run():
                _lib.check(lib.ffr_recnet_convlayer_fwd(P(x), n, cin, P(wp), cout, P(bias), P(slope), None, 0, 0, P(out),
                                                        cout, P(tab), 4, 81, None, None, _lib.stream_ptr()))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            row[name] = {"us": us, "tflops_algorithmic": 2.0 * n * 49 * cin * cout * 9 / us / 1e6}
        lib.ffr_debug_set_pixmajor(-1)
        print(json.dumps(row))
        rows.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/pix_conv_bench_%d.json" % n, "w"), indent=1)


if __name__ == "__main__":
    main()
