"""numpy model of the warp-level tensor-core fragment layouts the RecNet eval kernels rely on (csrc/recnet_kernels.cu:
recnet_prep_mma_kernel, feat_space_mma_kernel): `ldmatrix.m8n8.x4[.trans].b16` and `mma.sync.m16n8k16` (bf16) /
`mma.sync.m16n8k8` (tf32) as documented in the PTX ISA. Test infrastructure: it lets the index arithmetic of those kernels
be checked on the CPU (tests/test_mma_index_cpu.py restates the kernels' address formulas on top of this model)."""
import numpy as np


def ldmatrix_x4(arr, addr_of_lane, trans):
    """arr: flat array of 16-bit ELEMENTS; addr_of_lane(l) -> element index of the 8-element row lane l points to (lanes
    8i..8i+7 give the rows of matrix i). Returns regs[lane][i] = (lo, hi) element pair of matrix i held by the lane:
    plain: row lane/4, columns 2(lane%4), +1; .trans: rows 2(lane%4), +1 of column lane/4."""
    regs = [[None] * 4 for _ in range(32)]
    for mi in range(4):
        m = np.array([arr[addr_of_lane(8 * mi + r):addr_of_lane(8 * mi + r) + 8] for r in range(8)])
        for l in range(32):
            g, t = l // 4, l % 4
            regs[l][mi] = (m[g, 2 * t], m[g, 2 * t + 1]) if not trans else (m[2 * t, g], m[2 * t + 1, g])
    return regs


def mma_m16n8k16(acc, a, b0, b1):
    """acc[lane][4] += A(16x16, row) @ B(16x8, col). a[lane][4], b0[lane], b1[lane] are (lo, hi) pairs.
    A: a0 (g, 2t..), a1 (g+8, 2t..), a2 (g, 2t+8..), a3 (g+8, 2t+8..); B: b0 (k 2t.., n g), b1 (k 2t+8.., n g);
    C: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)."""
    A, B = np.zeros((16, 16)), np.zeros((16, 8))
    for l in range(32):
        g, t = l // 4, l % 4
        A[g, 2 * t], A[g, 2 * t + 1] = a[l][0]
        A[g + 8, 2 * t], A[g + 8, 2 * t + 1] = a[l][1]
        A[g, 2 * t + 8], A[g, 2 * t + 9] = a[l][2]
        A[g + 8, 2 * t + 8], A[g + 8, 2 * t + 9] = a[l][3]
        B[2 * t, g], B[2 * t + 1, g] = b0[l]
        B[2 * t + 8, g], B[2 * t + 9, g] = b1[l]
    _scatter_c(acc, A @ B)


def mma_m16n8k8(acc, a, b0, b1):
    """tf32: A(16x8): a0 (g, t), a1 (g+8, t), a2 (g, t+4), a3 (g+8, t+4); B(8x8): b0 (k t, n g), b1 (k t+4, n g)."""
    A, B = np.zeros((16, 8)), np.zeros((8, 8))
    for l in range(32):
        g, t = l // 4, l % 4
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[l]
        B[t, g], B[t + 4, g] = b0[l], b1[l]
    _scatter_c(acc, A @ B)


def _scatter_c(acc, C):
    for l in range(32):
        g, t = l // 4, l % 4
        acc[l][0] += C[g, 2 * t]
        acc[l][1] += C[g, 2 * t + 1]
        acc[l][2] += C[g + 8, 2 * t]
        acc[l][3] += C[g + 8, 2 * t + 1]


def new_acc(*dims):
    """nested lists [d0][d1]...[32 lanes][4] of zeros"""
    if not dims:
        return [[0.0] * 4 for _ in range(32)]
    return [new_acc(*dims[1:]) for _ in range(dims[0])]
