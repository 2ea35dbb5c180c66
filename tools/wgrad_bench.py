"""Developer timing script (GPU box): ffr_wgrad3x3 per RecNet layer shape at n samples (default 256). Not the contract bench."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib
from ffr_net_b200 import recnet_train as rt

SHAPES = [(561, 256), (256, 256), (256, 128), (128, 128), (128, 49), (49, 49), (1024, 512), (512, 512), (1536, 512)]
COUNT = {(561, 256): 1, (256, 256): 2, (256, 128): 1, (128, 128): 2, (128, 49): 1, (49, 49): 2, (1024, 512): 1,
         (512, 512): 4, (1536, 512): 1}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lib = _lib.load()
    P = n * 81
    out = {"n": n, "layers": []}
    total = 0.0
    for cin, cout in SHAPES:
        cin_p, cout_p = rt._ceil64(cin), rt._ceil64(cout)
        x = torch.randn(P, cin_p, device="cuda").bfloat16()
        dz = torch.randn(P, cout_p, device="cuda").bfloat16()
        dw = torch.empty(cout, cin, 3, 3, device="cuda")
        ws = torch.empty(9 * rt.wgrad_workspace_elems(cout, cin), device="cuda")

        def run():
            _lib.check(lib.ffr_wgrad3x3(_lib.ptr(dz), cout_p, _lib.ptr(x), cin_p, 0, n, cout, cin, _lib.ptr(dw), _lib.ptr(ws),
                                        _lib.stream_ptr()))
        row = {"cin": cin, "cout": cout}
        for splits in ([0] if (cin, cout) != (512, 512) else [0, 1, 2, 3, 4, 5, 8]):
            lib.ffr_debug_set_wgrad_splits(splits)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                run()
            e.record()
            torch.cuda.synchronize()
            us = s.elapsed_time(e) * 100
            tf = 2.0 * n * 49 * cin * cout * 9 / us / 1e6
            row["splits=%d" % splits] = {"us": us, "tflops_algorithmic": tf}
            if splits == 0:
                total += us * COUNT[(cin, cout)]
        lib.ffr_debug_set_wgrad_splits(0)
        print(json.dumps(row))
        out["layers"].append(row)
    out["total_us_per_recnet_backward"] = total
    print("sum over the 15 ConvLayers of one RecNet backward: %.1f us" % total)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/wgrad_bench_%d.json" % n, "w"), indent=1)


if __name__ == "__main__":
    main()
