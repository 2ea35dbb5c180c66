"""GPU parity of the fused CosFace head + CrossEntropy (ffr_net_b200/head.py, csrc/head_kernels.cu) against the fp32
CPU oracle (oracle.recnet.add_margin_product + F.cross_entropy, i.e. recnet.py:257-270 + trainer.py:173-176).
Tolerances: loss <= 2e-4 relative (hi/lo-split bf16 GEMM, ~1e-5 on the cosines); gradients <= 1e-2 relative L2
(dcos and the normalised operands of the two backward GEMMs are single bf16)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import recnet as orr
from ffr_net_b200 import _lib, head

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _case(n, classes, seed, aligned):
    g = torch.Generator().manual_seed(seed)
    w = (torch.rand(classes, 512, generator=g) * 2 - 1) * 0.0233
    v = torch.randn(n, 512, generator=g)
    label = torch.randint(0, classes, (n,), generator=g)
    if aligned:       # make some samples nearly parallel to their class vector: sharp softmax, cos near 1
        for i in range(0, n, 3):
            v[i] = w[label[i]] * 40 + 0.02 * torch.randn(512, generator=g)
    return w, v, label


@pytest.mark.parametrize("n,classes,aligned", [(5, 300, False), (64, 10575, True), (256, 10575, False), (130, 1000, True)])
def test_cosface_ce_forward_backward(lib, n, classes, aligned):
    w, v, label = _case(n, classes, 7 + n, aligned)
    wr, vr = w.clone().requires_grad_(True), v.clone().requires_grad_(True)
    logits, cos = orr.add_margin_product({"classifier.weight": wr}, vr, label)
    loss_ref = F.cross_entropy(logits, label)
    (loss_ref * 1.7).backward()
    wg, vg = w.cuda().requires_grad_(True), v.cuda().requires_grad_(True)
    loss, pred = head.cosface_ce(wg, vg, label.cuda())
    (loss * 1.7).backward()
    torch.cuda.synchronize()
    print("n=%d classes=%d loss %.6f ref %.6f | dv %.2e dW %.2e" %
          (n, classes, loss.item(), loss_ref.item(), rel_l2(vg.grad.cpu(), vr.grad), rel_l2(wg.grad.cpu(), wr.grad)))
    assert abs(loss.item() - loss_ref.item()) <= 2e-4 * max(1.0, abs(loss_ref.item()))
    # arg-max: identical wherever the top-2 cosine gap exceeds the GEMM error
    top2 = cos.detach().topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert torch.equal(pred.cpu()[clear], cos.detach().argmax(1)[clear])
    assert rel_l2(vg.grad.cpu(), vr.grad) <= 1e-2
    assert rel_l2(wg.grad.cpu(), wr.grad) <= 1e-2


def test_cosface_pack_hi_lo_split(lib):
    """ffr_cosface_pack: hi + lo reproduces the normalised fp32 row to ~2^-17, pad rows are zero, the transposed copy
    equals the hi part."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(70, 512, generator=g)
    xc = x.cuda()
    packed = torch.full((128, 1536), 7.0, dtype=torch.bfloat16, device="cuda")
    tr = torch.full((512, 128), 7.0, dtype=torch.bfloat16, device="cuda")
    for mode in (0, 1):
        _lib.check(lib.ffr_cosface_pack(_lib.ptr(xc), 70, 128, mode, _lib.ptr(packed), _lib.ptr(tr), 128, _lib.stream_ptr()))
        torch.cuda.synchronize()
        p = packed.float().cpu()
        hi = p[:, :512]
        lo = p[:, 512:1024] if mode == 0 else p[:, 1024:]
        dup = p[:, 1024:] if mode == 0 else p[:, 512:1024]
        ref = F.normalize(x)
        assert torch.equal(hi, dup)
        assert (hi[:70] + lo[:70] - ref).abs().max().item() <= 2e-6
        assert torch.equal(hi[:70], ref.bfloat16().float())
        assert p[70:].abs().max().item() == 0.0
        assert torch.equal(tr.float().cpu().t(), hi)
