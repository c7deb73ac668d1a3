"""Design validation for the round-2 head (DESIGN §6 "Next", item 2): cosine loss, labels and the gradient of the 17x17
score map WITHOUT materialising the (B, D, H, W) score tensor.  Pure torch on the CPU, small sizes, dense tap matrix —
this pins the algebra (tap weights with the crop at 19, border pixels with 1-2 taps, normalisation, gradient) that the
CUDA kernels will implement; `tests/test_fused_head_math.py` checks it against the oracle's materialised path.

Notation: S = (hs*ws, D) score map of one image, E = (C, D) class table, P = (H*W, hs*ws) bilinear tap matrix of
`upscore` + crop (models.py:94,146-147): up(S) = P S, at most 4 non-zeros per row.

    u_p . e_c      = (P (S E^T))[p, c]                    -> A = S E^T is hs*ws x C   (289 x 59 at full size)
    |u_p|^2        = sum_kl P[p,k] P[p,l] (S S^T)[k,l]    -> G = S S^T, neighbour entries only
    loss           = (N - sum_valid cos_p) / N,  cos_p = (P A)[p,t_p] / (|u_p| |e_t|)
    labels_p       = argmax_c (P A)[p,c] / |e_c|          (|e_c| == 0 -> 1; |u_p| is a common positive factor)
    dL/dS          = M1 E_hat + M2 S,  M1[k,c] = sum_{p: t_p = c} P[p,k] a_p,  M2[k,l] = sum_p P[p,k] b_p P[p,l]
                     a_p = -1 / (N |u_p|),  b_p = cos_p / (N |u_p|^2)
"""
import torch

CROP, KERNEL, STRIDE = 19, 64, 32


def tap_matrix(H, W, hs, ws, dtype=torch.float64):
    """Dense (H*W, hs*ws) matrix of the x32 bilinear transposed conv + crop 19."""
    def taps(n_out, n_in):
        m = torch.zeros(n_out, n_in, dtype=dtype)
        for y in range(n_out):
            yy = y + CROP
            for i in (yy // STRIDE, yy // STRIDE - 1):
                t = yy - STRIDE * i
                if 0 <= i < n_in and 0 <= t < KERNEL:
                    m[y, i] = 1.0 - abs(t - (KERNEL / 2 - 0.5)) / (KERNEL / 2)
        return m
    py, px = taps(H, hs), taps(W, ws)
    return torch.einsum("yi,xj->yxij", py, px).reshape(H * W, hs * ws)


def fused_cosine_head(s17, target, table, kind=0):
    """s17 (B, D, hs, ws), target (B, H, W) int64 with -1 = ignore, table (C, D).  kind 0 = cosine_loss, 1 = mse_loss
    (|u_p - e_t|^2 = |u_p|^2 - 2 u_p.e_t + |e_t|^2; a_p = -2/N on the raw rows, b_p = 2/N).
    Returns (loss, labels (B,H,W), d loss / d s17) computed from hs*ws-sized quantities and per-pixel scalars only."""
    B, D, hs, ws = s17.shape
    _, H, W = target.shape
    dt = torch.float64
    P = tap_matrix(H, W, hs, ws, dt)
    E = table.to(dt)
    en = E.norm(dim=1)
    en_fix = torch.where(en == 0, torch.ones_like(en), en)
    E_hat = E / en.clamp_min(1e-300)[:, None]  # target rows are never zero rows in practice; cos uses |e_t|
    n_valid = int((target >= 0).sum())
    total = torch.zeros((), dtype=dt)
    labels = torch.empty((B, H, W), dtype=torch.int64)
    grads = torch.zeros((B, hs * ws, D), dtype=dt)
    stats = []
    for b in range(B):
        S = s17[b].to(dt).reshape(D, hs * ws).t()          # (hs*ws, D)
        A = S @ E.t()                                       # (hs*ws, C)
        G = S @ S.t()                                       # (hs*ws, hs*ws)
        PA = P @ A                                          # per pixel: <= 4 taps x C
        un2 = ((P @ G) * P).sum(1)                          # |u_p|^2 from the Gram matrix
        un = un2.sqrt()
        labels[b] = (PA / en_fix[None]).argmax(1).view(H, W)
        t = target[b].reshape(-1)
        valid = t >= 0
        tc = t.clamp(min=0)
        pat = PA[torch.arange(H * W), tc]
        cos = pat / (un * en[tc]) if kind == 0 else un2 - 2.0 * pat + (en[tc] ** 2)
        total = total + cos[valid].sum()
        stats.append((P, S, un, cos, valid, tc))
    loss = (n_valid - total) / n_valid if kind == 0 else total / n_valid
    C = table.shape[0]
    for b, (P, S, un, cos, valid, tc) in enumerate(stats):
        if kind == 0:
            a = torch.where(valid, -1.0 / (n_valid * un), torch.zeros_like(un))
            bb = torch.where(valid, cos / (n_valid * un * un), torch.zeros_like(un))
        else:
            a = torch.where(valid, torch.full_like(un, -2.0 / n_valid), torch.zeros_like(un))
            bb = torch.where(valid, torch.full_like(un, 2.0 / n_valid), torch.zeros_like(un))
        onehot = torch.zeros(P.shape[0], C, dtype=dt)
        onehot[torch.arange(P.shape[0]), tc] = 1.0
        M1 = P.t() @ (onehot * a[:, None])                  # (hs*ws, C)
        M2 = P.t() @ (P * bb[:, None])                      # (hs*ws, hs*ws), neighbour-sparse
        grads[b] = M1 @ (E_hat if kind == 0 else E) + M2 @ S
    ds17 = grads.transpose(1, 2).reshape(B, D, hs, ws)
    return loss, labels, ds17
