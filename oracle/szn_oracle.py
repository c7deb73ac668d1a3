"""CPU oracle for the SZN pixel-embedding hot path (TEST INFRASTRUCTURE, not product code).

A plain ``torch`` (CPU, fp32) restatement of what the reference computes on the path
``FCN32s.forward`` -> loss -> ``infer_lbl*``.  Every function cites the reference lines it follows
(paths are relative to the reference checkout, commit 779fb72).

Parity status: the reference ships no tests and no golden vectors for this path, so the
pin is the reference *itself*: ``oracle/gen_golden.py`` imports the unmodified reference modules in the
dev container, runs them under seed 1337 and commits the results under ``tests/golden/``;
``tests/test_oracle_golden.py`` asserts this restatement reproduces them, and (when ``/root/reference``
is present) ``tests/test_oracle_vs_reference.py`` compares against the live reference.

The arithmetic itself lives in PyTorch (unpinned upstream; torch 2.11 CPU here).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# layer table of FCN32s.__init__ (models.py:39-100): name, Cin, Cout, kernel, padding
TRUNK = [
    ("conv1_1", 3, 64, 3, 100), ("conv1_2", 64, 64, 3, 1), ("pool1",),
    ("conv2_1", 64, 128, 3, 1), ("conv2_2", 128, 128, 3, 1), ("pool2",),
    ("conv3_1", 128, 256, 3, 1), ("conv3_2", 256, 256, 3, 1), ("conv3_3", 256, 256, 3, 1), ("pool3",),
    ("conv4_1", 256, 512, 3, 1), ("conv4_2", 512, 512, 3, 1), ("conv4_3", 512, 512, 3, 1), ("pool4",),
    ("conv5_1", 512, 512, 3, 1), ("conv5_2", 512, 512, 3, 1), ("conv5_3", 512, 512, 3, 1), ("pool5",),
]
CROP = 19          # models.py:147,151
UP_K, UP_S = 64, 32  # models.py:94,98


def bilinear_filter(k: int = UP_K) -> torch.Tensor:
    """The (k,k) tent filter of ``get_upsampling_weight`` (models.py:11-24), float64 -> float32."""
    f = (k + 1) // 2
    c = f - 1 if k % 2 == 1 else f - 0.5
    i = np.arange(k, dtype=np.float64)
    w1 = 1.0 - np.abs(i - c) / f
    return torch.from_numpy(np.outer(w1, w1)).float()


def upsampling_weight(cin: int, cout: int, k: int = UP_K) -> torch.Tensor:
    """Diagonal bilinear ConvTranspose2d weight (models.py:20-24)."""
    w = torch.zeros(cin, cout, k, k)
    filt = bilinear_filter(k)
    for i in range(min(cin, cout)):
        w[i, i] = filt
    return w


def init_params(n_class: int = 21, seed: int = 1337) -> dict:
    """Random parameters with the reference's names/shapes (models.py:39-112): torch default conv
    init for every conv (the zero-init is commented out upstream, models.py:104-108), bilinear deconvs."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def conv(name, cin, cout, k):
        fan_in = cin * k * k
        bound = 1.0 / np.sqrt(fan_in)
        p[name + ".weight"] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        p[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    for row in TRUNK:
        if len(row) == 5:
            conv(row[0], row[1], row[2], row[3])
    conv("fc6", 512, 4096, 7)
    conv("fc7", 4096, 4096, 1)
    conv("score_fr", 4096, n_class, 1)
    conv("seenmask_score", 4096, 2, 1)
    p["upscore.weight"] = upsampling_weight(n_class, n_class)
    p["seenmask_upscore.weight"] = upsampling_weight(2, 2)
    return p


def round_storage(t, storage):
    """Round an fp32 tensor to what the CUDA path stores: 'tf32' = cvt.rna.tf32.f32 (10-bit mantissa, nearest,
    ties away from zero), 'bf16' = round-to-nearest-even bfloat16; None = unchanged."""
    if storage is None:
        return t
    if storage == "bf16":
        return t.bfloat16().float()
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _RoundSTE(torch.autograd.Function):
    """y = round(x) in the forward pass (if ``fwd``), dx = round(dy) in the backward pass: the places where the
    CUDA path writes an activation / a data gradient to HBM in its storage type."""

    @staticmethod
    def forward(ctx, x, storage, fwd):
        ctx.storage = storage
        return round_storage(x, storage) if fwd else x.clone()

    @staticmethod
    def backward(ctx, g):
        return round_storage(g, ctx.storage), None, None


def _st(t, storage, fwd=True):
    return t if storage is None else _RoundSTE.apply(t, storage, fwd)


def _w(p, name, storage):
    """conv weight as the tensor-core kernels read it (rounded to the storage type; gradient passes through)."""
    w = p[name + ".weight"]
    return w if storage is None else w + (round_storage(w.detach(), storage) - w.detach())


def trunk_forward(x, p, drop_masks=None, collect=None, storage=None):
    """conv1_1 .. drop7 of FCN32s.forward (models.py:115-143).  ``drop_masks`` = (m6, m7) of shape
    (B,4096) with values in {0,1}; when given they are applied as Dropout2d(p=.5) would (x * m * 2),
    otherwise dropout is the identity (eval mode).

    ``storage`` ('tf32' | 'bf16' | None) makes this a *storage-precision emulation* of the CUDA path: the same
    fp32 arithmetic, but every tensor the kernels keep in HBM in a narrower type (packed weights, activations, data
    gradients) is rounded at the same point.  With it, ReLU / max-pool decisions agree with the GPU, so gradients can
    be compared tightly; ``storage=None`` is the reference (fp32) semantics."""
    h = x
    for row in TRUNK:
        if len(row) == 1:
            h = F.max_pool2d(h, 2, stride=2, ceil_mode=True)
        else:
            name, _, _, _, pad = row
            w = p[name + ".weight"] if name == "conv1_1" else _w(p, name, storage)  # conv1_1 runs in fp32 FMA
            h = _st(F.relu(F.conv2d(h, w, p[name + ".bias"], padding=pad)), storage)
        if collect is not None:
            collect[row[0]] = h
    h = F.relu(F.conv2d(h, _w(p, "fc6", storage), p["fc6.bias"]))
    if drop_masks is not None:
        h = h * (drop_masks[0][:, :, None, None] * 2.0)
    h = _st(h, storage)
    if collect is not None:
        collect["fc6"] = h
    h = F.relu(F.conv2d(h, _w(p, "fc7", storage), p["fc7.bias"]))
    if drop_masks is not None:
        h = h * (drop_masks[1][:, :, None, None] * 2.0)
    h = _st(h, storage)
    if collect is not None:
        collect["fc7"] = h
    return h


def head_forward(h, p, score_name, up_name, H, W, storage=None):
    """score conv -> x32 transposed conv -> crop (models.py:145-151)."""
    s = F.conv2d(h, _w(p, score_name, storage), p[score_name + ".bias"])
    s = _st(s, storage, fwd=False)  # the 17x17 score map stays fp32; its gradient is stored in the narrow type
    s = F.conv_transpose2d(s, p[up_name + ".weight"], stride=UP_S)
    return s[:, :, CROP:CROP + H, CROP:CROP + W].contiguous()


def forward(x, p, mode="fcn", drop_masks=None, collect=None, storage=None):
    """FCN32s.forward (models.py:114-160): both heads always evaluated, selection by ``mode``."""
    H, W = x.shape[2], x.shape[3]
    h = trunk_forward(x, p, drop_masks, collect, storage)
    f = head_forward(h, p, "score_fr", "upscore", H, W, storage)
    s = head_forward(h, p, "seenmask_score", "seenmask_upscore", H, W, storage)
    if mode == "fcn":
        return f
    if mode == "seenmask":
        return s
    if mode == "both":
        return f, s
    raise Exception("model given unexpected forward mode")


# ---------------------------------------------------------------------------------------------
# losses (utils.py:19-102).  The reference's cosine_loss/infer_lbl are only correct for n == 1
# (SURVEY §0.4); the batched meaning used here is the n == 1 formula applied with keepdim norms and
# the valid-pixel count taken over the whole batch, which for n == 1 is the reference bit for bit.
# ---------------------------------------------------------------------------------------------

def cross_entropy2d(score, target, size_average=False):
    """utils.py:19-48: log_softmax over c, pick target class where target >= 0, sum; /N_valid if asked."""
    logp = F.log_softmax(score, dim=1)
    valid = target >= 0
    t = target.clamp(min=0)
    picked = logp.gather(1, t[:, None]).squeeze(1)
    loss = -(picked * valid).sum()
    if size_average:
        loss = loss / valid.sum()
    return loss


def mse_loss(score, target, target_embed):
    """utils.py:50-73: sum over valid pixels and channels of (s-e)^2, divided by the pixel count."""
    valid = (target >= 0)[:, None]
    d = (score - target_embed) * valid
    return (d * d).sum() / valid.sum()


def cosine_loss(score, target, target_embed):
    """utils.py:75-102: (N - sum_valid cos(s, e)) / N with both vectors L2-normalised over c."""
    sn = score / score.norm(p=2, dim=1, keepdim=True)
    en = target_embed / target_embed.norm(p=2, dim=1, keepdim=True)
    valid = target >= 0
    n = valid.sum()
    cos = (sn * en).sum(1)
    # masked_select then sum, like the reference's boolean-mask gather (ignored pixels may hold NaN)
    return (n - cos[valid].sum()) / n


def target_embed_from_labels(target, table):
    """Dataset-side gather (pascal_dataset.py:122-128, context_dataset.py:128-133): E[label], with
    label -1 mapped to class 0; (n,h,w) int64 -> (n,D,h,w) fp32."""
    t = target.clamp(min=0)
    return table[t].permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------------------------------------
# inference (utils.py:159-205)
# ---------------------------------------------------------------------------------------------

def infer_lbl(score, embed_arr):
    """utils.py:159-185 per sample: sim = S E^T / (|s| * |e| with |e|==0 -> 1); arg-max over classes
    (first index on ties, torch CPU max).  Returns (n,h,w) int64 numpy."""
    n, c, h, w = score.shape
    out = []
    for i in range(n):
        s = score[i].permute(1, 2, 0).reshape(h * w, c)
        sim = s @ embed_arr.t()
        sn = s.norm(p=2, dim=1, keepdim=True)
        en = embed_arr.norm(p=2, dim=1)[None, :].clone()
        en[en == 0] = 1
        sim = sim / (sn * en)
        out.append(sim.max(1)[1].view(h, w))
    return torch.stack(out).numpy()


def split_tables(table, unseen):
    """trainer_fcn.py:44,55-64: seen / unseen copies of the table with the other rows zeroed."""
    C = table.shape[0]
    seen = [c for c in range(C) if c not in unseen]
    se, ue = torch.zeros_like(table), torch.zeros_like(table)
    se[seen] = table[seen]
    ue[list(unseen)] = table[list(unseen)]
    return se, ue


def stitch(score, seen_tab, unseen_tab, unseen_mask):
    """utils.py:201-205."""
    pred = infer_lbl(score, seen_tab)
    alt = infer_lbl(score, unseen_tab)
    pred[unseen_mask] = alt[unseen_mask]
    return pred


def infer_lbl_forced_unseen(score, target, seen_tab, unseen_tab, unseen):
    """utils.py:188-192: unseen mask from the ground truth."""
    m = np.isin(target.numpy(), list(unseen))
    return stitch(score, seen_tab, unseen_tab, m)


def infer_lbl_szn(score, seen_mask_score, seen_tab, unseen_tab):
    """utils.py:195-199: unseen mask = 1 - argmax of the 2-channel seenmask score."""
    m = (1 - seen_mask_score.max(1)[1].numpy()).astype(bool)
    return stitch(score, seen_tab, unseen_tab, m)


def seenmask_target(target, unseen, n_class):
    """trainer_seenmask.py:55-56: 1 where the label is a seen class, else 0 (so -1 -> 0)."""
    seen = [c for c in range(n_class) if c not in unseen]
    return torch.from_numpy(np.isin(target.numpy(), seen).astype(np.int64))


# ---------------------------------------------------------------------------------------------
# metrics (utils.py:104-154), the "next" row of SURVEY §8f
# ---------------------------------------------------------------------------------------------

def fast_hist(lt, lp, n_class):
    """utils.py:104-121 (target='all')."""
    m = (lt >= 0) & (lt < n_class)
    return np.bincount(n_class * lt[m].astype(int) + lp[m], minlength=n_class ** 2).reshape(n_class, n_class)


def hist_to_metrics(hist):
    """utils.py:123-131."""
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        mean_iu = np.nanmean(iu)
        freq = hist.sum(axis=1) / hist.sum()
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
    return acc, acc_cls, mean_iu, fwavacc


# ---------------------------------------------------------------------------------------------
# synthetic workload shared by tests and bench
# ---------------------------------------------------------------------------------------------
MEAN_BGR = (104.00698793, 116.66876762, 122.67891434)  # pascal_dataset.py:39


def synth_batch(B, H, W, C, D, seed=1337, block=32, ignore_frac=0.05):
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
    x = img - torch.tensor(MEAN_BGR).view(1, 3, 1, 1)
    hb, wb = (H + block - 1) // block, (W + block - 1) // block
    lab = torch.randint(0, C, (B, hb, wb), generator=g)
    lab = lab.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()
    ign = torch.rand(B, H, W, generator=g) < ignore_frac
    lab[ign] = -1
    table = torch.randn(C, D, generator=g)
    table = table / table.norm(dim=1).max()
    return x, lab, table
