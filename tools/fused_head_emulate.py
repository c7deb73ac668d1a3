"""Index-for-index Python emulation of csrc/szn_fused_head.cu (nodes / pixels / grad kernels: same cell and tap numbering,
same Gram-pair table, same M2 expansion, same gather of the <= 4 cells around a node), in float64.  It exists because the
kernels were written without GPU access: `tests/test_fused_head_math.py` checks this emulation against the dense-matrix
algebra of tools/fused_head_math.py, so an indexing slip in the kernel design shows up on the CPU."""
import numpy as np


def nodes(s17, table):
    """s17 (B, hs, ws, D); returns A (B,K,C), G (B,K,5): self, right, down, down-right, down-left."""
    B, hs, ws, D = s17.shape
    K = hs * ws
    A = np.zeros((B, K, table.shape[0]))
    G = np.zeros((B, K, 5))
    for b in range(B):
        for k in range(K):
            i, j = divmod(k, ws)
            me = s17[b, i, j]
            A[b, k] = table @ me
            for q, (ni, nj) in enumerate(((i, j), (i, j + 1), (i + 1, j), (i + 1, j + 1), (i + 1, j - 1))):
                ok = ni < hs and 0 <= nj < ws
                G[b, k, q] = me @ s17[b, ni, nj] if ok else 0.0
    return A, G


def pixels(A, G, en_inv, en2, kind, target, H, W, hs, ws):
    """Per cell (anchor (ia, ja), grid (hs+1) x (ws+1)): labels, sum of cos, count, M1cell (4,C), M2cell (4,4)."""
    B, K, C = A.shape
    cells = (hs + 1) * (ws + 1)
    labels = np.zeros((B, H, W), dtype=np.int64)
    M1 = np.zeros((B, cells, 4, C))
    M2 = np.zeros((B, cells, 4, 4))
    tot, cnt = 0.0, 0.0
    src = (0, 1, 2, 3, 0, 2, 0, 1, 0, 1)
    qq = (0, 0, 0, 0, 1, 1, 2, 2, 3, 4)
    pair_of = {(0, 1): 4, (2, 3): 5, (0, 2): 6, (1, 3): 7, (0, 3): 8, (1, 2): 9}
    for b in range(B):
        for by in range(hs + 1):
            for bx in range(ws + 1):
                ia, ja = by - 1, bx - 1
                cell = by * (ws + 1) + bx
                node, ok = [0] * 4, [False] * 4
                for a in range(4):
                    ni, nj = ia + (a >> 1), ja + (a & 1)
                    ok[a] = 0 <= ni < hs and 0 <= nj < ws
                    node[a] = ni * ws + nj if ok[a] else 0
                Ac = np.stack([A[b, node[a]] if ok[a] else np.zeros(C) for a in range(4)])
                Gc = [G[b, node[src[e]], qq[e]] if ok[src[e]] else 0.0 for e in range(10)]
                m2 = np.zeros(10)
                for ty in range(32):
                    y = 32 * (ia + 1) - 19 + ty
                    if not 0 <= y < H:
                        continue
                    wy = ((31.5 - ty) / 32, (ty + 0.5) / 32)
                    for tx in range(32):
                        x = 32 * (ja + 1) - 19 + tx
                        if not 0 <= x < W:
                            continue
                        wx = ((31.5 - tx) / 32, (tx + 0.5) / 32)
                        w = np.array([wy[0] * wx[0], wy[0] * wx[1], wy[1] * wx[0], wy[1] * wx[1]])
                        pa = w @ Ac
                        labels[b, y, x] = int(np.argmax(pa * en_inv))  # first maximum
                        t = int(target[b, y, x])
                        if 0 <= t < C:
                            un2 = sum(w[a] * w[a] * Gc[a] for a in range(4)) + 2 * sum(
                                w[a] * w[a2] * Gc[e] for (a, a2), e in pair_of.items())
                            if kind == 0:
                                inv_un = 1.0 / np.sqrt(un2)
                                cs = pa[t] * inv_un * en_inv[t]
                                tot += cs
                                ap, bp = -inv_un, cs * inv_un * inv_un
                            else:
                                tot += un2 - 2.0 * pa[t] + en2[t]
                                ap, bp = -2.0, 2.0
                            cnt += 1
                            M1[b, cell, :, t] += w * ap
                            for a in range(4):
                                m2[a] += w[a] * w[a] * bp
                            for (a, a2), e in pair_of.items():
                                m2[e] += w[a] * w[a2] * bp
                for a in range(4):
                    for a2 in range(4):
                        lo, hi = min(a, a2), max(a, a2)
                        M2[b, cell, a, a2] = m2[lo] if lo == hi else m2[pair_of[(lo, hi)]]
    return labels, tot, cnt, M1, M2


def grad(s17, table, en_inv, kind, M1, M2, n_valid, gout=1.0):
    B, hs, ws, D = s17.shape
    ds = np.zeros_like(s17)
    for b in range(B):
        for k in range(hs * ws):
            i, j = divmod(k, ws)
            m1n = np.zeros(table.shape[0])
            acc = np.zeros(D)
            for a in range(4):
                ia, ja = i - (a >> 1), j - (a & 1)
                cell = (ia + 1) * (ws + 1) + (ja + 1)
                m1n += M1[b, cell, a]
                for a2 in range(4):
                    ni, nj = ia + (a2 >> 1), ja + (a2 & 1)
                    if 0 <= ni < hs and 0 <= nj < ws:
                        acc += M2[b, cell, a, a2] * s17[b, ni, nj]
            acc += (m1n * en_inv if kind == 0 else m1n) @ table
            ds[b, i, j] = acc * (gout / n_valid)
    return ds


def fused_cosine_head(s17_nchw, target, table, kind=0):
    """Same interface as tools/fused_head_math.fused_cosine_head (numpy in / out)."""
    s17 = np.ascontiguousarray(np.transpose(s17_nchw, (0, 2, 3, 1))).astype(np.float64)
    table = table.astype(np.float64)
    B, hs, ws, D = s17.shape
    _, H, W = target.shape
    en = np.sqrt((table * table).sum(1))
    en_inv = np.where(en == 0, 1.0, 1.0 / np.where(en == 0, 1.0, en))
    A, G = nodes(s17, table)
    labels, tot, cnt, M1, M2 = pixels(A, G, en_inv, en * en, kind, target, H, W, hs, ws)
    loss = (cnt - tot) / cnt if kind == 0 else tot / cnt
    ds = grad(s17, table, en_inv, kind, M1, M2, cnt)
    return loss, labels, np.transpose(ds, (0, 3, 1, 2))
