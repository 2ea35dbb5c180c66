import subprocess, sys, os
code = '''
import sys, torch
sys.path.insert(0, %r)
from ffr_net_b200 import _lib
lib = _lib.load()
variant, r0 = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(1)
a = torch.randint(-3, 4, (96, 128), generator=g, device="cuda").to(torch.bfloat16)
b = torch.randint(-3, 4, (96, 64), generator=g, device="cuda").to(torch.bfloat16)
out = torch.zeros(128, 64, dtype=torch.float32, device="cuda")
_lib.check(_lib.load_probe().ffr_debug_mn_probe(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), r0, variant, _lib.stream_ptr()))
torch.cuda.synchronize()
ref = a[:64].float().t() @ b[r0:r0 + 64].float()
print("variant", variant, "r0", r0, "equal", bool(torch.equal(out, ref)), "maxdiff", (out-ref).abs().max().item(), "nz", (out!=0).float().mean().item())
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for variant in (0, 1):
    for r0 in (0, 8, 1, 9, 10):
        p = subprocess.run([sys.executable, "-c", code, str(variant), str(r0)], capture_output=True, text=True, timeout=120)
        print((p.stdout.strip() or p.stderr.strip()[-300:]))
