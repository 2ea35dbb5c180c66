"""Torch-side views of the device layouts (used by tests and by one-off conversions, never per batch on the hot path).

Halo-shared flat NHWC: an (N,C,S,S) map is an (N*(S+1)^2, C) row-major matrix; pixel (n,h,w) sits in row
n*(S+1)^2 + h*(S+1) + w and rows with h == S or w == S are zero. The zero row/column is shared: it is the right
neighbour of column S-1, the left neighbour of column 0 of the next row, the bottom neighbour of row S-1 and the top
neighbour of row 0 of the next image, so a 3x3/pad-1 tap is a pure row offset (r-1)*(S+1) + (s-1).
"""
import torch


def to_flat(x, dtype=torch.bfloat16):
    n, c, s, _ = x.shape
    out = torch.zeros(n, s + 1, s + 1, c, dtype=dtype, device=x.device)
    out[:, :s, :s, :] = x.permute(0, 2, 3, 1).to(dtype)
    return out.reshape(n * (s + 1) * (s + 1), c)


def from_flat(t, n, s, c):
    return t.reshape(n, s + 1, s + 1, c)[:, :s, :s, :].permute(0, 3, 1, 2).float().contiguous()


def flat_pad_rows(t, n, s, c):
    """The pad rows/columns of a flat map (must be all zero)."""
    v = t.reshape(n, s + 1, s + 1, c)
    return torch.cat([v[:, s, :, :].reshape(-1), v[:, :, s, :].reshape(-1)])


def to_s2d(x, dtype=torch.bfloat16):
    """(N,C,S,S) -> space-to-depth flat map: rows of the (S/2+1)^2 grid, channels [(h%2)*2 + (w%2)]*C + c."""
    n, c, s, _ = x.shape
    so = s // 2
    v = x.permute(0, 2, 3, 1).reshape(n, so, 2, so, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(n, so, so, 4 * c)
    out = torch.zeros(n, so + 1, so + 1, 4 * c, dtype=dtype, device=x.device)
    out[:, :so, :so, :] = v.to(dtype)
    return out.reshape(n * (so + 1) * (so + 1), 4 * c)


def from_s2d(t, n, s, c):
    so = s // 2
    v = t.reshape(n, so + 1, so + 1, 2, 2, c)[:, :so, :so]
    return v.permute(0, 5, 1, 3, 2, 4).reshape(n, c, s, s).float().contiguous()
