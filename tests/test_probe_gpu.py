"""Hardware-semantics probe: row-offset UMMA descriptors on a SWIZZLE_128B tile (csrc/probe.cu).
Records which descriptor variant addresses rows [r0, r0+128) correctly; the sliding-window conv kernel relies on it."""
import json
import os

import pytest
import torch

from ffr_net_b200 import _lib

pytestmark = pytest.mark.gpu


def test_rowshift_descriptor_semantics(lib):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randint(-4, 5, (256, 64), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randint(-2, 3, (64, 64), generator=g, device="cuda").to(torch.bfloat16)
    result = {}
    for variant in (0, 1):
        ok = []
        for r0 in (0, 1, 3, 7, 8, 15, 16, 17, 29, 64, 100, 128):
            out = torch.zeros(128, 64, dtype=torch.float32, device="cuda")
            _lib.check(_lib.load_probe().ffr_debug_rowshift_probe(_lib.ptr(a), _lib.ptr(w), _lib.ptr(out), r0, variant,
                                                    _lib.stream_ptr()))
            torch.cuda.synchronize()
            ref = a[r0:r0 + 128].float() @ w.float().t()
            ok.append(bool(torch.equal(out, ref)))
        result["variant%d" % variant] = ok
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_rowshift.json", "w") as f:
        json.dump(result, f)
    print("rowshift probe:", result)
    assert result["variant0"][0] and result["variant1"][0], "aligned descriptor must work in both variants"
    assert all(result["variant0"]), result


def test_mn_major_descriptor_semantics(lib):
    """MN-major SWIZZLE_128B operands (needed by the weight-gradient GEMM, which contracts over pixel rows)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(-3, 4, (96, 128), generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randint(-3, 4, (96, 64), generator=g, device="cuda").to(torch.bfloat16)
    result = {}
    for variant in (0,):     # variant 1 (LBO/SBO swapped) faults the context on B200 — established once, not re-run
        ok = []
        for r0 in (0, 8, 16, 1, 3, 9, 10, 19, 32):
            out = torch.zeros(128, 64, dtype=torch.float32, device="cuda")
            _lib.check(_lib.load_probe().ffr_debug_mn_probe(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), r0, variant, _lib.stream_ptr()))
            torch.cuda.synchronize()
            ref = a[:64].float().t() @ b[r0:r0 + 64].float()
            ok.append(bool(torch.equal(out, ref)))
        result["variant%d" % variant] = ok
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_mn_major.json", "w") as f:
        json.dump(result, f)
    print("mn-major probe:", result)
    assert all(result["variant0"]), result


def test_fp16_operands(lib):
    """tcgen05.mma kind::f16 with fp16 operands (a_format = b_format = F16): the training forward GEMMs use fp16
    activations (hi + lo) and fp16 weights. MIXED formats (fp16 x bf16 in one instruction) raise an illegal-instruction
    fault on B200 (measured once, profiles/r02_probe_mixed_formats.json) — which is why the weight-gradient GEMM contracts
    the bf16 dz with a bf16 COPY of the activations."""
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randint(-3, 4, (96, 128), generator=g, device="cuda").float()
    b = torch.randint(-3, 4, (96, 64), generator=g, device="cuda").float()
    a = a + 0.0009765625 * torch.randint(0, 2, (96, 128), generator=g, device="cuda")   # 2^-10: exact in fp16, not in bf16
    aa, bb = a.half(), b.half()
    out = torch.zeros(128, 64, dtype=torch.float32, device="cuda")
    _lib.check(_lib.load_probe().ffr_debug_mn_probe(_lib.ptr(aa), _lib.ptr(bb), _lib.ptr(out), 0, 6, _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = aa[:64].float().t() @ bb[:64].float()
    err = float((out - ref).abs().max())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_f16.json", "w") as f:
        json.dump({"f16xf16_max_abs_err": err}, f)
    assert err == 0.0
