"""Test helpers (TEST ONLY): torch restatements of the H9 layout (7x7 map with its reflection halo, rows n*81 + hp*9 + wp)
and of the fp16 hi/lo split, used to build kernel inputs and expected values. Never imported by the product."""
import torch


def own_rows():
    """Indices (within the 81 rows of one image) of the 49 interior pixels, in pixel order h*7+w."""
    return torch.tensor([(h + 1) * 9 + (w + 1) for h in range(7) for w in range(7)])


def to_h9(x, cpad=None, mirror=True):
    """(n,C,7,7) -> [n*81, cpad] fp32: reflection halo filled (mirror) or zero."""
    n, c = x.shape[0], x.shape[1]
    cpad = cpad or c
    if mirror:
        p = torch.nn.functional.pad(x.float(), (1, 1, 1, 1), mode="reflect")
    else:
        p = torch.nn.functional.pad(x.float(), (1, 1, 1, 1))
    out = torch.zeros(n, 81, cpad, dtype=torch.float32, device=x.device)
    out[:, :, :c] = p.permute(0, 2, 3, 1).reshape(n, 81, c)
    return out.reshape(n * 81, cpad)


def own_to_h9(x, cpad=None):
    """(n,C,7,7) -> [n*81, cpad] fp32 with the values on the own rows only (halo rows zero)."""
    return to_h9(x, cpad, mirror=False)


def from_h9(rows, c, n=None):
    """[n*81, ld] -> (n,c,7,7) from the own rows."""
    ld = rows.shape[1]
    n = n or rows.shape[0] // 81
    r = rows.reshape(n, 81, ld)[:, own_rows().to(rows.device), :c]
    return r.reshape(n, 7, 7, c).permute(0, 3, 1, 2).contiguous()


def fold_h9(rows, c, n=None):
    """Gradient fold of the reflection fan-out: [n*81, ld] gradient on the padded grid -> (n,c,7,7) (sum over the
    positions that mirror each interior pixel) — the adjoint of to_h9(mirror=True)."""
    ld = rows.shape[1]
    n = n or rows.shape[0] // 81
    g = rows.reshape(n, 9, 9, ld)[..., :c].permute(0, 3, 1, 2).double()
    x = torch.zeros(n, c, 7, 7, dtype=torch.float64, device=rows.device, requires_grad=True)
    p = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="reflect")
    (p * g).sum().backward()
    return x.grad.float()


def hilo(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return hi, lo


def hilo_cat(x):
    """[rows, C] fp32 -> [rows, 2C] fp16 = [hi | lo]."""
    hi, lo = hilo(x)
    return torch.cat((hi, lo), 1).contiguous()


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
