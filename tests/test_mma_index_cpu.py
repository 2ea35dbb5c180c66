"""CPU check of the index arithmetic of the warp-MMA RecNet eval kernels (csrc/recnet_kernels.cu): the ldmatrix lane
addresses, the tile -> warp assignment, the accumulator -> (row, column) maps and the K-slot permutation that feeds the
tf32 maps from the previous accumulators are restated here line by line on top of a numpy model of the fragment layouts
(tests/mma_model.py) and compared in float64 with the plain matrix products. The formulas below mirror the kernels
(same pitches, same expressions); this is how the kernels were validated before their first GPU run."""
import numpy as np

from mma_model import ldmatrix_x4, mma_m16n8k16, mma_m16n8k8, new_acc

XP, BP, WP, AP, MP = 520, 40, 72, 36, 72           # PM_XP, PM_BP, PM_WP, PM_AP, FS_MP


def _setup():
    rng = np.random.default_rng(0)
    d = dict(x=rng.standard_normal((512, 49)), w0aT=rng.standard_normal((49, 32)), w0bT=rng.standard_normal((512, 32)),
             b0=rng.standard_normal(32), c1=rng.standard_normal(32), c2=rng.standard_normal(32),
             A1=rng.standard_normal((32, 32)) * 0.2, A2=rng.standard_normal((32, 32)) * 0.2,
             sl1=rng.uniform(0.1, 0.4, 512), sl4=rng.uniform(0.1, 0.4, 512), sl7=rng.uniform(0.1, 0.4, 512),
             M=rng.uniform(0, 1, (49, 49)))
    x = d["x"]
    xt = np.zeros(64 * XP)
    for hw in range(49):
        xt[hw * XP:hw * XP + 512] = x[:, hw]
    d["xt"] = xt
    d["inv_c"] = 1 / np.sqrt((x ** 2).sum(1))
    inv_s = np.zeros(64)
    inv_s[:49] = 1 / np.sqrt((x ** 2).sum(0))
    d["inv_s"] = inv_s
    return d


def test_recnet_prep_mma_index_logic():
    d = _setup()
    x, xt, inv_c, inv_s = d["x"], d["xt"], d["inv_c"], d["inv_s"]
    Wc = np.zeros(64 * WP)
    for hw in range(49):
        Wc[hw * WP:hw * WP + 32] = d["w0aT"][hw]
    Bs = np.zeros(512 * BP)
    for c in range(512):
        Bs[c * BP:c * BP + 32] = d["w0bT"][c] * inv_c[c]
    A1s, A2s = np.zeros(32 * AP), np.zeros(32 * AP)
    for i in range(32):
        A1s[i * AP:i * AP + 32], A2s[i * AP:i * AP + 32] = d["A1"][i], d["A2"][i]
    # ---- spatial Gram, warps 0-7 ----
    Gs = np.zeros((49, 49))
    for warp in range(8):
        mt, ng = warp >> 1, warp & 1
        acc = new_acc(4)
        a_addr = lambda l: (16 * mt + (l & 7) + ((l >> 3) & 1) * 8) * XP + (l >> 4) * 8
        b_addr = lambda l: (32 * ng + (l & 7) + (l >> 4) * 8) * XP + ((l >> 3) & 1) * 8
        for ks in range(32):
            a = ldmatrix_x4(xt, lambda l: a_addr(l) + ks * 16, False)
            b01 = ldmatrix_x4(xt, lambda l: b_addr(l) + ks * 16, False)
            b23 = ldmatrix_x4(xt, lambda l: b_addr(l) + 16 * XP + ks * 16, False)
            for q, (b, lo) in enumerate(((b01, 0), (b01, 2), (b23, 0), (b23, 2))):
                mma_m16n8k16(acc[q], a, [r[lo] for r in b], [r[lo + 1] for r in b])
        for q in range(4):
            for l in range(32):
                g, t = l // 4, l % 4
                for e in range(4):
                    i, j = 16 * mt + g + (e >> 1) * 8, 32 * ng + 8 * q + 2 * t + (e & 1)
                    if i < 49 and j < 49:
                        Gs[i, j] = acc[q][l][e] * inv_s[i] * inv_s[j]
    xn = x / np.sqrt((x ** 2).sum(0, keepdims=True))
    assert np.abs(Gs - xn.T @ xn).max() < 1e-12
    # ---- T = Xh^T W0b^T, warps 8-15, written into the second half of [W0a | T] ----
    for w8 in range(8):
        mt, nh = w8 >> 1, w8 & 1
        acc = new_acc(2)
        a_addr = lambda l: (16 * mt + (l & 7) + ((l >> 3) & 1) * 8) * XP + (l >> 4) * 8
        b_addr = lambda l: ((l & 7) + ((l >> 3) & 1) * 8) * BP + 16 * nh + (l >> 4) * 8
        for ks in range(32):
            a = ldmatrix_x4(xt, lambda l: a_addr(l) + ks * 16, False)
            b = ldmatrix_x4(Bs, lambda l: b_addr(l) + ks * 16 * BP, True)
            mma_m16n8k16(acc[0], a, [r[0] for r in b], [r[1] for r in b])
            mma_m16n8k16(acc[1], a, [r[2] for r in b], [r[3] for r in b])
        for q in range(2):
            for l in range(32):
                g, t = l // 4, l % 4
                for e in (0, 2):
                    hw, j = 16 * mt + g + (e >> 1) * 8, 16 * nh + 8 * q + 2 * t
                    Wc[hw * WP + 32 + j], Wc[hw * WP + 32 + j + 1] = acc[q][l][e], acc[q][l][e + 1]
    T_ref = (x * inv_c[:, None]).T @ d["w0bT"]
    T_got = np.array([Wc[hw * WP + 32:hw * WP + 64] for hw in range(64)])
    assert np.abs(T_got[:49] - T_ref).max() < 1e-11 and np.abs(T_got[49:]).max() == 0.0
    # ---- chain: bf16 first layer, two tf32 maps fed from the accumulators ----
    misc = np.concatenate([d["b0"], d["c1"], d["c2"]])
    H5 = np.zeros((512, 32))
    for warp in range(16):
        m_base = 32 * warp
        acc = new_acc(2, 8)
        a_addr = lambda l: ((l & 7) + (l >> 4) * 8) * XP + m_base + ((l >> 3) & 1) * 8
        b_addr = lambda l: ((l & 7) + ((l >> 3) & 1) * 8) * WP + (l >> 4) * 8
        for ks in range(4):
            a0 = ldmatrix_x4(xt, lambda l: a_addr(l) + ks * 16 * XP, True)
            a1 = ldmatrix_x4(xt, lambda l: a_addr(l) + ks * 16 * XP + 16, True)
            for pq in range(4):
                b = ldmatrix_x4(Wc, lambda l: b_addr(l) + ks * 16 * WP + pq * 16, True)
                for m, a in ((0, a0), (1, a1)):
                    mma_m16n8k16(acc[m][2 * pq], a, [r[0] for r in b], [r[1] for r in b])
                    mma_m16n8k16(acc[m][2 * pq + 1], a, [r[2] for r in b], [r[3] for r in b])
        h = new_acc(2, 4)
        for l in range(32):
            g, t = l // 4, l % 4
            for m in range(2):
                for r in range(2):
                    c = m_base + 16 * m + g + 8 * r
                    for q in range(4):
                        for e2 in range(2):
                            e = 2 * r + e2
                            v = inv_c[c] * acc[m][q + 4][l][e] + (acc[m][q][l][e] + misc[8 * q + 2 * t + e2])
                            h[m][q][l][e] = v if v > 0 else v * d["sl1"][c]

        def chain_map(As, cvec, slope):
            o = new_acc(2, 4)
            for l in range(32):
                t = l % 4
                for q in range(4):
                    for m in range(2):
                        o[m][q][l][0] = o[m][q][l][2] = cvec[8 * q + 2 * t]
                        o[m][q][l][1] = o[m][q][l][3] = cvec[8 * q + 2 * t + 1]
            for s2 in range(4):
                a = [[(h[m][s2][l][0], h[m][s2][l][2], h[m][s2][l][1], h[m][s2][l][3]) for l in range(32)] for m in range(2)]
                for q in range(4):
                    b0 = [As[(8 * q + l // 4) * AP + 8 * s2 + 2 * (l % 4)] for l in range(32)]
                    b1 = [As[(8 * q + l // 4) * AP + 8 * s2 + 2 * (l % 4) + 1] for l in range(32)]
                    mma_m16n8k8(o[0][q], a[0], b0, b1)
                    mma_m16n8k8(o[1][q], a[1], b0, b1)
            for l in range(32):
                g = l // 4
                for m in range(2):
                    for q in range(4):
                        for e in range(4):
                            v = o[m][q][l][e]
                            h[m][q][l][e] = v if v > 0 else v * slope[m_base + 16 * m + g + 8 * (e >> 1)]
        chain_map(A1s, misc[32:64], d["sl4"])
        chain_map(A2s, misc[64:96], d["sl7"])
        for l in range(32):
            g, t = l // 4, l % 4
            for m in range(2):
                for r in range(2):
                    c = m_base + 16 * m + g + 8 * r
                    for q in range(4):
                        H5[c, 8 * q + 2 * t], H5[c, 8 * q + 2 * t + 1] = h[m][q][l][2 * r], h[m][q][l][2 * r + 1]
    prelu = lambda v, s: np.where(v > 0, v, v * s[:, None])
    h1 = prelu(x @ d["w0aT"] + d["b0"] + inv_c[:, None] * (x @ T_ref), d["sl1"])
    h4 = prelu(h1 @ d["A1"].T + d["c1"], d["sl4"])
    h7 = prelu(h4 @ d["A2"].T + d["c2"], d["sl7"])
    assert np.abs(H5 - h7).max() < 1e-11


def test_feat_space_mma_index_logic():
    d = _setup()
    x, xt, M = d["x"], d["xt"], d["M"]
    Ms = np.zeros(64 * MP)
    for i in range(49):
        Ms[i * MP:i * MP + 49] = M[i]
    out = np.zeros((49, 512))
    for warp in range(8):
        n_base = 64 * warp
        for mp in range(2):
            acc = new_acc(2, 8)
            for ks in range(4):
                a = [ldmatrix_x4(Ms, lambda l, m=m: (16 * ks + (l & 7) + (l >> 4) * 8) * MP + 32 * mp + 16 * m + ((l >> 3) & 1) * 8, True)
                     for m in range(2)]
                for pq in range(4):
                    b = ldmatrix_x4(xt, lambda l: (16 * ks + (l & 7) + ((l >> 3) & 1) * 8) * XP + n_base + 16 * pq + (l >> 4) * 8, True)
                    for m in range(2):
                        mma_m16n8k16(acc[m][2 * pq], a[m], [r[0] for r in b], [r[1] for r in b])
                        mma_m16n8k16(acc[m][2 * pq + 1], a[m], [r[2] for r in b], [r[3] for r in b])
            for m in range(2):
                for q in range(8):
                    for l in range(32):
                        g, t = l // 4, l % 4
                        for e in range(4):
                            j, c = 16 * (2 * mp + m) + g + (e >> 1) * 8, n_base + 8 * q + 2 * t + (e & 1)
                            if j < 49:
                                out[j, c] = acc[m][q][l][e]
    assert np.abs(out - (x @ M).T).max() < 1e-11
