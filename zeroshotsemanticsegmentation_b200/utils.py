"""Drop-in for the hot-path functions of the reference's ``utils.py``: same names, argument meaning
and return types (``utils.py:19-102`` losses, ``utils.py:159-205`` inference), executed by fused
sm_100a kernels.  Differences, all supersets of the reference behaviour:

* ``cosine_loss`` / ``infer_lbl*`` accept any batch size (the reference is only correct for n == 1,
  SURVEY §0.4); for n == 1 the result is the reference's.
* the embedding losses also accept ``target_embed=None, table=E``: the per-pixel target vector is then
  gathered from the (C, D) class table on the device (what the datasets do on the host,
  ``pascal_dataset.py:122-128``), which avoids materialising / uploading the (n, D, h, w) target.
* ``accum_hook`` (a callable applied in place to the device accumulator ``[sum, n_valid]`` before the loss is
  formed, e.g. ``torch.distributed.all_reduce``) lets data-parallel training normalise by the GLOBAL valid count.

There is no CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import pickle

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr


def load_obj(name):
    """utils.py:11-13."""
    with open(name + ".pkl", "rb") as f:
        return pickle.load(f, encoding="latin-1")


def save_obj(obj, name):
    """utils.py:15-17."""
    with open(name + ".pkl", "wb") as f:
        pickle.dump(obj, f, pickle.HIGHEST_PROTOCOL)


def split_embeddings(embed_arr, unseen):
    """Seen / unseen copies of the (C, D) class table with the other set's rows zeroed, as ``trainer_fcn.Trainer``
    builds them (``trainer_fcn.py:44,55-64``).  A zero row scores exactly 0 in ``infer_lbl`` (``utils.py:175``)."""
    table = torch.as_tensor(embed_arr).float()
    unseen = sorted(set(int(u) for u in unseen))
    seen = [c for c in range(table.shape[0]) if c not in unseen]
    seen_t, unseen_t = torch.zeros_like(table), torch.zeros_like(table)
    seen_t[seen] = table[seen]
    unseen_t[unseen] = table[unseen]
    return seen_t, unseen_t


def seenmask_target(target, unseen, n_class):
    """Binary target of the seen-mask phase (``trainer_seenmask.py:53-56``): 1 where the label is a seen class, else 0
    (the ignore label -1 therefore becomes 0 and is NOT ignored, as upstream).  Labels below -1 are batch padding
    (``trainer.collate_padded``), which upstream never has: they stay negative, so the loss and the metrics skip them.
    Works on the tensor's own device, so the label map need not visit the host."""
    unseen = torch.as_tensor(sorted(set(int(u) for u in unseen)), dtype=torch.long, device=target.device)
    t = target.long()
    seen = (t >= 0) & (t < n_class) & ~torch.isin(t, unseen)
    return torch.where(t < -1, t, seen.long())


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("zeroshotsemanticsegmentation_b200.utils runs on CUDA tensors only (no CPU fallback)")


def _as_f32(t):
    return t.detach().contiguous().float() if t is not None else None


def _check_shapes(score, target, target_embed=None, table=None):
    """The kernels index raw memory: reject shapes that would read out of bounds before anything is launched."""
    if score.dim() != 4:
        raise ValueError("score must be (n, c, h, w), got %s" % (tuple(score.shape),))
    n, c, h, w = score.shape
    if target is not None and tuple(target.shape) != (n, h, w):
        raise ValueError("target must be (n, h, w) = %s, got %s" % ((n, h, w), tuple(target.shape)))
    if target_embed is not None and tuple(target_embed.shape) != (n, c, h, w):
        raise ValueError("target_embed must have the score's shape %s, got %s" % ((n, c, h, w), tuple(target_embed.shape)))
    if table is not None and (table.dim() != 2 or table.shape[1] != c):
        raise ValueError("the class table must be (C, %d), got %s" % (c, tuple(table.shape)))


class _EmbedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, target_embed, table, kind, accum_hook):
        _check_cuda(score, target, target_embed, table)
        _check_shapes(score, target, target_embed, table)
        if target_embed is None and table is None:
            raise ValueError("give either target_embed (n,c,h,w) or table (C,c) to gather it from")
        n, c, h, w = score.shape
        sc = _as_f32(score)
        tg = target.detach().contiguous().long()
        te, tb = _as_f32(target_embed), _as_f32(table)
        dev = score.device
        stats = torch.empty((n * h * w * 3,), device=dev, dtype=torch.float32) if kind == 0 else None
        accum = torch.empty(2, device=dev, dtype=torch.float64)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        st = _lib.stream()
        rows = 0 if tb is None else tb.shape[0]  # labels >= rows poison the loss with NaN instead of reading past the table
        call("szn_embed_loss_fwd", kind, ptr(sc), ptr(tg), ptr(te), ptr(tb), rows, n, c, h, w, ptr(stats), ptr(accum),
             ptr(loss), st)
        if accum_hook is not None:  # e.g. all-reduce of {sum, n_valid} across data-parallel ranks
            accum_hook(accum)
            call("szn_loss_finalize", kind, ptr(accum), ptr(loss), st)
        ctx.saved = (sc, tg, te, tb, stats, accum, kind)
        return loss

    @staticmethod
    def backward(ctx, gout):
        sc, tg, te, tb, stats, accum, kind = ctx.saved
        n, c, h, w = sc.shape
        g = torch.empty_like(sc)
        go = gout.detach().contiguous().float()
        call("szn_embed_loss_bwd", kind, ptr(sc), ptr(tg), ptr(te), ptr(tb), 0 if tb is None else tb.shape[0], n, c, h, w,
             ptr(stats), ptr(accum), ptr(go), ptr(g), _lib.stream())
        return g, None, None, None, None, None


def _fused_handle(score):
    """The ScoreHandle of a score returned by ``FCN32s(fused_head=True)`` if it still describes ``score``, else None."""
    head = getattr(score, "_szn_head", None)
    return head if head is not None and head.valid_for(score) else None


def _fused_workspace(s17, C):
    B, hs, ws, _ = s17.shape
    n = int(_lib.load().szn_head_fused_workspace_floats(B, hs, ws, C))
    return torch.empty(n, device=s17.device, dtype=torch.float32)


class _FusedHeadLoss(torch.autograd.Function):
    """EXPERIMENTAL: ``cosine_loss`` (kind 0) / ``mse_loss`` (kind 1) on the hs x ws score map (``szn_head_fused_*``);
    input and gradient are [B,hs,ws,Dp] fp32, the (B, D, H, W) score tensor is not touched."""

    @staticmethod
    def forward(ctx, s17, target, table, D, hw, accum_hook, kind=0, head=None):
        _check_cuda(s17, target, table)
        B, hs, ws, ld = s17.shape
        H, W = hw
        sc = s17.detach().contiguous().float()
        tg = target.detach().contiguous().long()
        tb = _as_f32(table)
        if tb.shape[1] != D or tuple(tg.shape) != (B, H, W):
            raise ValueError("fused head: table / target do not match the score")
        C = tb.shape[0]
        work = _fused_workspace(sc, C)
        accum = torch.empty(2, device=sc.device, dtype=torch.float64)
        loss = torch.empty((), device=sc.device, dtype=torch.float32)
        st = _lib.stream()
        # the loss pass scans every class of every pixel anyway: it also writes the nearest-embedding labels, which
        # infer_lbl* on the same score and table then simply picks up (trainer_fcn.py:100-117 calls both every iteration)
        labels = torch.empty((B, H, W), device=sc.device, dtype=torch.int64) if head is not None else None
        call("szn_head_fused_fwd", kind, ptr(sc), ld, 0, ptr(tg), ptr(tb), B, D, H, W, hs, ws, C, ptr(work), ptr(accum),
             ptr(loss), ptr(labels), st)
        if head is not None:
            head.labels_cache = (table.data_ptr(), table._version, tuple(table.shape), labels)
        if accum_hook is not None:
            accum_hook(accum)
            call("szn_loss_finalize", kind, ptr(accum), ptr(loss), st)
        ctx.saved = (sc, tb, work, accum, D, kind)
        return loss

    @staticmethod
    def backward(ctx, gout):
        sc, tb, work, accum, D, kind = ctx.saved
        B, hs, ws, ld = sc.shape
        ds = torch.empty_like(sc)
        go = gout.detach().contiguous().float()
        call("szn_head_fused_bwd", kind, ptr(sc), ld, 0, ptr(tb), B, D, hs, ws, tb.shape[0], ptr(work), ptr(accum), ptr(go),
             ptr(ds), _lib.stream())
        return ds, None, None, None, None, None, None, None


class _CrossEntropy2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, size_average, accum_hook):
        _check_cuda(score, target)
        _check_shapes(score, target)
        n, c, h, w = score.shape
        sc = _as_f32(score)
        tg = target.detach().contiguous().long()
        dev = score.device
        lse = torch.empty((n * h * w,), device=dev, dtype=torch.float32)
        accum = torch.empty(2, device=dev, dtype=torch.float64)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        st = _lib.stream()
        call("szn_ce2d_fwd", ptr(sc), ptr(tg), n, c, h, w, int(bool(size_average)), ptr(lse), ptr(accum), ptr(loss), st)
        if accum_hook is not None:  # data parallel: {sum, n_valid} over ALL ranks (the mean divides by the global count)
            accum_hook(accum)
            call("szn_loss_finalize", 3 if size_average else 2, ptr(accum), ptr(loss), st)
        ctx.saved = (sc, tg, lse, accum, int(bool(size_average)))
        return loss

    @staticmethod
    def backward(ctx, gout):
        sc, tg, lse, accum, sa = ctx.saved
        n, c, h, w = sc.shape
        g = torch.empty_like(sc)
        go = gout.detach().contiguous().float()
        call("szn_ce2d_bwd", ptr(sc), ptr(tg), n, c, h, w, sa, ptr(lse), ptr(accum), ptr(go), ptr(g), _lib.stream())
        return g, None, None, None


def cross_entropy2d(score, target, weight=None, size_average=False, accum_hook=None):
    """Per-pixel softmax cross entropy, summed over ``target >= 0`` (``utils.py:19-48``).
    score (n,c,h,w), target (n,h,w) int64; ``size_average`` divides by the number of valid pixels."""
    if weight is not None:
        raise NotImplementedError("class weights are never passed by the reference trainers")
    return _CrossEntropy2d.apply(score, target, size_average, accum_hook)


def mse_loss(score, target, target_embed=None, table=None, accum_hook=None):
    """sum over valid pixels and channels of (score - target_embed)^2 / n_valid (``utils.py:50-73``)."""
    head = _fused_handle(score) if target_embed is None and table is not None else None
    if head is not None:  # FCN32s(fused_head=True)
        return _FusedHeadLoss.apply(head.s17, target, table, head.D, head.hw, accum_hook, 1, head)
    return _EmbedLoss.apply(score, target, target_embed, table, 1, accum_hook)


def cosine_loss(score, target, target_embed=None, table=None, accum_hook=None):
    """(N - sum_valid cos(score_p, target_embed_p)) / N (``utils.py:75-102``)."""
    head = _fused_handle(score) if target_embed is None and table is not None else None
    if head is not None:  # FCN32s(fused_head=True): work from the 17x17 map the score was upsampled from
        return _FusedHeadLoss.apply(head.s17, target, table, head.D, head.hw, accum_hook, 0, head)
    return _EmbedLoss.apply(score, target, target_embed, table, 0, accum_hook)


def _labels_device(score, embed_arr):
    _check_cuda(score, embed_arr)
    _check_shapes(score, None, None, embed_arr)
    n, c, h, w = score.shape
    head = _fused_handle(score)
    if head is not None:  # FCN32s(fused_head=True)
        cached = getattr(head, "labels_cache", None)
        if cached is not None and cached[:3] == (embed_arr.data_ptr(), embed_arr._version, tuple(embed_arr.shape)):
            return cached[3]  # written by the loss pass on this score and this table
        s17 = head.s17.detach().contiguous().float()
        tb = _as_f32(embed_arr)
        if tb.shape[1] != c:
            raise ValueError("embedding width %d does not match score channels %d" % (tb.shape[1], c))
        out = torch.empty((n, h, w), device=score.device, dtype=torch.int64)
        work = _fused_workspace(s17, tb.shape[0])
        call("szn_head_fused_fwd", 0, ptr(s17), s17.shape[3], 0, None, ptr(tb), n, c, h, w, s17.shape[1], s17.shape[2],
             tb.shape[0], ptr(work), None, None, ptr(out), _lib.stream())
        return out
    sc, tb = _as_f32(score), _as_f32(embed_arr)
    C, D = tb.shape
    if D != c:
        raise ValueError("embedding width %d does not match score channels %d" % (D, c))
    en = torch.empty(int(_lib.load().szn_embed_argmax_scratch_floats(C, D)), device=score.device, dtype=torch.float32)
    out = torch.empty((n, h, w), device=score.device, dtype=torch.int64)
    call("szn_embed_argmax", ptr(sc), ptr(tb), n, c, h, w, C, ptr(en), ptr(out), _lib.stream())
    return out


def infer_lbl_device(score, embed_arr):
    """``infer_lbl`` that leaves the (n,h,w) int64 labels on the device (no D2H, no numpy)."""
    return _labels_device(score, embed_arr)


def infer_lbl(score, embed_arr, cuda=True):
    """Nearest class embedding by cosine similarity, zero rows score 0 (``utils.py:159-185``).
    Returns an ``np.ndarray`` (n,h,w) int64 like the reference."""
    return _labels_device(score, embed_arr).cpu().numpy()


def _stitch(score, seen_embed_arr, unseen_embed_arr, seen_mask_score=None, target=None, unseen=None):
    seen_lbl = _labels_device(score, seen_embed_arr)
    unseen_lbl = _labels_device(score, unseen_embed_arr)
    n, h, w = seen_lbl.shape
    if seen_mask_score is not None and tuple(seen_mask_score.shape) != (n, 2, h, w):
        raise ValueError("seen_mask_score must be (n, 2, h, w) = %s, got %s" % ((n, 2, h, w), tuple(seen_mask_score.shape)))
    if target is not None and tuple(target.shape) != (n, h, w):
        raise ValueError("target must be (n, h, w) = %s, got %s" % ((n, h, w), tuple(target.shape)))
    out = torch.empty_like(seen_lbl)
    sm = _as_f32(seen_mask_score)
    tg = target.detach().contiguous().long() if target is not None else None
    un = torch.as_tensor(list(unseen), device=score.device, dtype=torch.int64) if unseen is not None else None
    call("szn_stitch_labels", ptr(seen_lbl), ptr(unseen_lbl), ptr(sm), ptr(tg), ptr(un),
         0 if un is None else un.numel(), n, h, w, ptr(out), _lib.stream())
    return out


def infer_lbl_forced_unseen(score, target, seen_embed_arr, unseen_embed_arr, unseen, cuda=True):
    """``utils.py:188-192``: pixels whose ground truth is an unseen class are labelled among unseen classes only."""
    _check_cuda(target)
    return _stitch(score, seen_embed_arr, unseen_embed_arr, target=target, unseen=unseen).cpu().numpy()


def infer_lbl_szn(score, seen_mask_score, seen_embed_arr, unseen_embed_arr, cuda=True):
    """``utils.py:195-199``: the seen-mask head decides per pixel which table is used."""
    _check_cuda(seen_mask_score)
    return _stitch(score, seen_embed_arr, unseen_embed_arr, seen_mask_score=seen_mask_score).cpu().numpy()


def infer_lbl_forced_unseen_device(score, target, seen_embed_arr, unseen_embed_arr, unseen):
    """``infer_lbl_forced_unseen`` that leaves the labels on the device (int64 CUDA tensor (n,h,w))."""
    _check_cuda(target)
    return _stitch(score, seen_embed_arr, unseen_embed_arr, target=target, unseen=unseen)


def infer_lbl_szn_device(score, seen_mask_score, seen_embed_arr, unseen_embed_arr):
    """``infer_lbl_szn`` that leaves the labels on the device (int64 CUDA tensor (n,h,w))."""
    _check_cuda(seen_mask_score)
    return _stitch(score, seen_embed_arr, unseen_embed_arr, seen_mask_score=seen_mask_score)


def stich_seen_unseen_with_mask(score, seen_embed_arr, unseen_embed_arr, unseen_mask, cuda=True):
    """``utils.py:201-205`` with an explicit boolean mask (n,h,w)."""
    pred = infer_lbl(score, seen_embed_arr)
    alt = infer_lbl(score, unseen_embed_arr)
    pred[unseen_mask] = alt[unseen_mask]
    return pred


# ---- host-side metrics (utils.py:104-154); device version is SURVEY §8f row 1 ----

def _fast_hist(label_true, label_pred, n_class, target="all", unseen=None):
    keep = (label_true >= 0) & (label_true < n_class)
    if target == "unseen":
        keep &= np.isin(label_true, unseen)
    elif target == "seen":
        keep &= np.isin(label_true, [c for c in range(n_class) if c not in unseen])
    idx = n_class * label_true[keep].astype(int) + label_pred[keep]
    return np.bincount(idx, minlength=n_class ** 2).reshape(n_class, n_class)


def _hist_to_metrics(hist):
    with np.errstate(divide="ignore", invalid="ignore"):
        tp = np.diag(hist)
        acc = tp.sum() / hist.sum()
        acc_cls = np.nanmean(tp / hist.sum(axis=1))
        iu = tp / (hist.sum(axis=1) + hist.sum(axis=0) - tp)
        freq = hist.sum(axis=1) / hist.sum()
        return acc, acc_cls, np.nanmean(iu), (freq[freq > 0] * iu[freq > 0]).sum()


def metrics_from_hist(hist):
    """(k, n_class, n_class) confusion matrices (device tensor or array) -> k tuples (acc, acc_cls, mean_iu, fwavacc),
    ``utils.py:123-131``.  Lets a validation loop accumulate one histogram on the device instead of every label map."""
    if isinstance(hist, torch.Tensor):
        hist = hist.cpu().numpy()
    hist = np.asarray(hist, dtype=np.float64)
    if hist.ndim == 2:
        hist = hist[None]
    return tuple(_hist_to_metrics(h) for h in hist)


def confusion_hist_device(label_true, label_pred, n_class, unseen=None):
    """Confusion matrices of ``_fast_hist`` (``utils.py:104-121``) for CUDA label tensors, built on the device.
    Returns an int64 CUDA tensor (1 or 3, n_class, n_class): target 'all' [, 'seen', 'unseen']."""
    _check_cuda(label_true, label_pred)
    if label_true.numel() != label_pred.numel():
        raise ValueError("label_true and label_pred differ in size: %s vs %s" % (tuple(label_true.shape), tuple(label_pred.shape)))
    lt = label_true.detach().contiguous().long().view(-1)
    lp = label_pred.detach().contiguous().long().view(-1)
    flags = None
    if unseen:
        flags = torch.zeros(n_class, dtype=torch.uint8, device=lt.device)
        flags[torch.as_tensor(list(unseen), device=lt.device, dtype=torch.long)] = 1
    hist = torch.zeros((3 if unseen else 1, n_class, n_class), dtype=torch.int64, device=lt.device)
    call("szn_confusion_hist", ptr(lt), ptr(lp), lt.numel(), n_class, ptr(flags), ptr(hist), _lib.stream())
    return hist


def label_accuracy_score(label_trues, label_preds, n_class, unseen=None):
    """Pixel accuracy, mean class accuracy, mean IU, frequency-weighted IU (``utils.py:133-154``).
    Accepts the reference's numpy label maps, or CUDA tensors (a tensor or a sequence of tensors): then the histograms
    are built on the device and only n_class^2 counters are copied to the host."""
    first = label_trues[0] if isinstance(label_trues, (list, tuple)) else label_trues
    if isinstance(first, torch.Tensor) and first.is_cuda:
        lts = label_trues if isinstance(label_trues, (list, tuple)) else [label_trues]
        lps = label_preds if isinstance(label_preds, (list, tuple)) else [label_preds]
        hist = sum(confusion_hist_device(a, b, n_class, unseen) for a, b in zip(lts, lps)).cpu().numpy().astype(np.float64)
        res = [_hist_to_metrics(h) for h in hist]
        return res[0] if not unseen else tuple(res)
    kinds = ["all"] + (["seen", "unseen"] if unseen else [])
    hists = {k: np.zeros((n_class, n_class)) for k in kinds}
    for lt, lp in zip(label_trues, label_preds):
        for k in kinds:
            hists[k] += _fast_hist(lt.flatten(), lp.flatten(), n_class, target=k, unseen=unseen)
    res = [_hist_to_metrics(hists[k]) for k in kinds]
    return res[0] if not unseen else tuple(res)
