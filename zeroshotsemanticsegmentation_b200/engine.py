"""Device-side execution of ``FCN32s.forward`` and its backward through ``libszn.so``.

Follows the reference layer sequence (``models.py:114-160``) and what autograd does for it
(``trainer_fcn.py:157``), but every device op is one of the hand-written kernels behind the C ABI
(``include/szn.h``).  PyTorch supplies device memory, the current stream and the autograd hook only.

Data layout in HBM
  * image ``x``: NCHW fp32 (public API, read once by ``szn_conv1_1_fwd``);
  * trunk activations: NHWC, element type = precision (``tf32``: fp32 rounded to TF32, ``bf16``);
  * packed weights in the activation type, cached per parameter version: ``[Cout][R*S][Cin]`` for the forward pass,
    ``[Cin][R*S flipped][Cout]`` for the data gradient (fc6: ``[R*S*Cin][Cout]`` + ``szn_col2im``);
  * conv parameters live in channels_last memory, so the wgrad kernel's ``[Cout][R][S][Cin]`` fp32 output buffer is
    handed to autograd as ``weight.grad`` without a transposing pass;
  * head: ``s17`` = fp32 ``[B,hs,ws,Dp]`` holding ``score_fr`` (channels 0..D-1) and ``seenmask_score``
    (channels D, D+1) from ONE GEMM, ``Dp`` = D+2 rounded up to 64;
  * returned ``f`` (B,D,H,W) / ``s`` (B,2,H,W): NCHW fp32 contiguous, as the reference returns them.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr

# (name, Cin, Cout, kernel, pad) or ("poolN",): models.py:43-81
TRUNK = [
    ("conv1_1", 3, 64, 3, 100), ("conv1_2", 64, 64, 3, 1), ("pool1",),
    ("conv2_1", 64, 128, 3, 1), ("conv2_2", 128, 128, 3, 1), ("pool2",),
    ("conv3_1", 128, 256, 3, 1), ("conv3_2", 256, 256, 3, 1), ("conv3_3", 256, 256, 3, 1), ("pool3",),
    ("conv4_1", 256, 512, 3, 1), ("conv4_2", 512, 512, 3, 1), ("conv4_3", 512, 512, 3, 1), ("pool4",),
    ("conv5_1", 512, 512, 3, 1), ("conv5_2", 512, 512, 3, 1), ("conv5_3", 512, 512, 3, 1), ("pool5",),
]
CONV_NAMES = [r[0] for r in TRUNK if len(r) == 5] + ["fc6", "fc7"]
# order of the parameter tensors handed to the autograd function
PARAM_ORDER = [n + s for n in CONV_NAMES for s in (".weight", ".bias")] + [
    "score_fr.weight", "score_fr.bias", "seenmask_score.weight", "seenmask_score.bias",
    "upscore.weight", "seenmask_upscore.weight"]

# precision -> (C-ABI dtype code, torch dtype of the trunk tensors).  "fp32" is the fp32-grade mode: every activation /
# data gradient / packed weight is stored as two bf16 planes (hi, lo) -- a pixel row of C channels is a bf16 row of 2C
# elements, the 4C bytes an fp32 row would take -- and every product is hi*hi + lo*hi + hi*lo on the bf16 tensor cores.
PRECISIONS = {"tf32": (_lib.F32, torch.float32), "bf16": (_lib.BF16, torch.bfloat16),
              "fp32": (_lib.F32X3, torch.bfloat16)}


def planes(dt):
    """bf16 planes per stored value (2 for the split fp32-grade format)."""
    return 2 if dt == _lib.F32X3 else 1


import os as _os
# SZN_POOL_CODE=1: one routing byte per pooled element (szn_pool_fwd_code / szn_pool_bwd_code) instead of keeping and
# re-reading the pre-pool activation in the max-pool backward; 0: the y-reading pair.  New in this build.
_NEW_KERNELS_DEFAULT = "1"
_POOL_Y = _os.environ.get("SZN_POOL_CODE", _NEW_KERNELS_DEFAULT) != "1"
_NO_FUSED_DB = _os.environ.get("SZN_NO_FUSED_DB") == "1"  # A/B switch: bias gradients through szn_bias_grad passes instead of the dgrad / pool_bwd epilogues


def round_up(a, b):
    return (a + b - 1) // b * b


class PackedWeights:
    """Per-module cache of kernel-layout weights, refreshed when a parameter's version changes.  Edits through ``.data``
    (``p.data.copy_()``, the style the reference's own init / copy code uses) do NOT advance the version counter: call
    ``invalidate()`` after them -- ``FCN32s`` does so itself in ``copy_params_from_vgg16``, ``load_state_dict`` and
    ``_apply`` (``.to()`` / ``.cuda()`` / ``.float()``)."""

    def __init__(self):
        self.cache = {}

    def invalidate(self):
        self.cache.clear()

    def get(self, key, version, build):
        hit = self.cache.get(key)
        if hit is not None and hit[0] == version:
            return hit[1]
        val = build()
        self.cache[key] = (version, val)
        return val


def _pack(dt, tdtype, w, o_pad=None):
    w = w.contiguous()  # parameters may be channels_last (models._use_kernel_weight_layout); the pack kernels read OIHW
    O, I, R, S = w.shape
    o_pad = O if o_pad is None else o_pad
    out = torch.empty((planes(dt) * o_pad, R * S, I), device=w.device, dtype=tdtype)  # split: hi plane, then lo plane
    call("szn_pack_weight", dt, ptr(w), ptr(out), O, I, R, S, o_pad, _lib.stream())
    return out


def _pack_d(dt, tdtype, w, mode=0, o_pad=None):
    """Transposed weights for the data gradient: [Cin][R*S flipped][O_pad] (mode 0) or [R*S*Cin][O_pad] (mode 1)."""
    w = w.contiguous()
    O, I, R, S = w.shape
    o_pad = O if o_pad is None else o_pad
    out = torch.empty((planes(dt) * I, R * S, o_pad), device=w.device, dtype=tdtype)
    call("szn_pack_weight_dgrad", dt, ptr(w), ptr(out), O, I, R, S, o_pad, mode, _lib.stream())
    return out


def is_diag_bilinear(w):
    """True when ``w`` (Ci,Co,64,64) is exactly what ``_initialize_weights`` wrote (models.py:102-112)."""
    from .models import bilinear_filter
    if w.shape[0] != w.shape[1] or w.shape[2] != 64 or w.shape[3] != 64:
        return False
    n = w.shape[0]
    idx = torch.arange(n, device=w.device)
    diag = w[idx, idx]
    filt = bilinear_filter(64).to(w.device)
    if not torch.equal(diag, filt.expand_as(diag)):
        return False
    return int(torch.count_nonzero(w)) == int(torch.count_nonzero(diag))


class ScoreHandle:
    """Attached to a score returned by ``FCN32s(fused_head=True)``: the hs x ws map ``s17`` ([B,hs,ws,Dp] fp32, an autograd
    output of the same node) the score was upsampled from.  ``valid_for`` is false once the score has been modified in
    place or is another tensor (``detach()``, slicing, arithmetic make new tensors without the handle)."""

    def __init__(self, s17, score):
        self.s17 = s17
        self.D = score.shape[1]
        self.hw = (score.shape[2], score.shape[3])
        self._version = score._version
        self._ptr = score.data_ptr()

    def valid_for(self, score):
        return score._version == self._version and score.data_ptr() == self._ptr and tuple(score.shape[2:]) == self.hw


def _finish(module, grads):
    if module._grad_flush is not None:
        module._grad_flush()
    return (None, None) + tuple(grads[n] for n in PARAM_ORDER)


class _GradDict(dict):
    """name -> gradient; tells ``on_ready(name, tensor)`` (the data-parallel reducer) the moment a gradient is final,
    so its all-reduce overlaps the rest of the backward pass."""

    def __init__(self, on_ready):
        super().__init__()
        self.on_ready = on_ready

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, v)
        if v is not None and self.on_ready is not None:
            self.on_ready(k, v)


class FCN32sFunction(torch.autograd.Function):
    """(x, *params) -> (f, s); both heads are always evaluated (models.py:145-151)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        ctx.set_materialize_grads(False)  # an unused head (mode='fcn' / 'seenmask') arrives as None, not as zeros
        dt, tdtype = PRECISIONS[module.precision]
        npl = planes(dt)
        st = _lib.stream()
        dev = x.device
        if x.dtype != torch.float32:
            raise TypeError("FCN32s expects an fp32 image batch")
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("FCN32s expects a (B, 3, H, W) image batch, got %s" % (tuple(x.shape),))
        if x.requires_grad:
            raise NotImplementedError("the gradient with respect to the image is not computed (conv1_1 has no dgrad)")
        for name, prm in zip(PARAM_ORDER, params):
            # the kernels read raw fp32 storage on the input's device: anything else would be read as garbage
            if prm.dtype != torch.float32 or prm.device != dev:
                raise TypeError("parameter %s is %s on %s; the B200 path needs fp32 parameters on the input's device (%s)"
                                % (name, prm.dtype, prm.device, dev))
        x = x.contiguous()
        B, _, H, W = x.shape
        P = dict(zip(PARAM_ORDER, params))
        D = P["score_fr.weight"].shape[0]
        Dp = round_up(D + 2, 64)
        pw = module._packed

        def packed(name, o_pad=None):
            w = P[name + ".weight"]
            return pw.get((name, dt), (w._version, w.data_ptr()), lambda: _pack(dt, tdtype, w.detach(), o_pad))

        acts = {}   # name -> NHWC activation (post-ReLU / pooled)
        dims = {}   # name -> (H, W, C) of that activation
        codes = {}  # pool name -> routing codes of szn_pool_fwd_code
        # inference (no_grad) and head-only training (frozen trunk: trainer_seenmask.py) never run a pool backward
        trunk_bwd = getattr(module, "_grad_enabled", True) and any(
            g for n, g in zip(PARAM_ORDER, ctx.needs_input_grad[2:]) if n.startswith("conv"))
        # conv1_1 on CUDA cores straight from the NCHW image
        h, w_ = H + 198, W + 198
        a = torch.empty((B, h, w_, npl * 64), device=dev, dtype=tdtype)
        call("szn_conv1_1_fwd", dt, ptr(x), ptr(P["conv1_1.weight"].detach().contiguous()),
             ptr(P["conv1_1.bias"].detach()), ptr(a), B, H, W, 100, st)
        acts["conv1_1"], dims["conv1_1"] = a, (h, w_, 64)
        c = 64
        prev_name = "conv1_1"
        for row in TRUNK[1:]:
            name = row[0]
            if len(row) == 1:
                ho, wo = (h + 1) // 2, (w_ + 1) // 2
                o = torch.empty((B, ho, wo, npl * c), device=dev, dtype=tdtype)
                if _POOL_Y or not trunk_bwd:
                    call("szn_pool_fwd", dt, ptr(a), ptr(o), B, h, w_, c, st)
                else:
                    # winner + ReLU-gate byte per pooled element: all the backward pass needs of the pre-pool activation,
                    # which is dropped here (conv1_2's output alone is 1 GB at B = 8, 512 x 512)
                    codes[name] = torch.empty((B, ho, wo, c), device=dev, dtype=torch.uint8)
                    call("szn_pool_fwd_code", dt, ptr(a), ptr(o), ptr(codes[name]), B, h, w_, c, st)
                    acts[prev_name] = None
                h, w_ = ho, wo
            else:
                _, cin, cout, k, pad = row
                o = torch.empty((B, h, w_, npl * cout), device=dev, dtype=tdtype)
                with _lib.nvtx_range(name + " fwd"):
                    call("szn_conv_fwd", dt, ptr(a), ptr(packed(name)), ptr(P[name + ".bias"].detach()), ptr(o),
                         B, h, w_, cin, cout, k, k, pad, 1, None, 0, 0, cout, st)
                c = cout
            a = o
            acts[name], dims[name] = a, (h, w_, c)
            prev_name = name
        # fc6 (7x7 valid) / fc7 (1x1) with ReLU and Dropout2d folded into the epilogue
        training = module.training
        drop = None
        if training:
            drop = torch.empty((2, B, 4096), device=dev, dtype=torch.float32)
            if module._forced_drop_masks is not None:
                m6, m7 = module._forced_drop_masks
                drop[0].copy_(m6.to(dev).float() * 2.0)
                drop[1].copy_(m7.to(dev).float() * 2.0)
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
                call("szn_dropout_scale", ptr(drop), 2 * B * 4096, seed, st)
        hs, ws = h - 6, w_ - 6
        h6 = torch.empty((B, hs, ws, npl * 4096), device=dev, dtype=tdtype)
        call("szn_conv_fwd", dt, ptr(a), ptr(packed("fc6")), ptr(P["fc6.bias"].detach()), ptr(h6), B, h, w_, 512, 4096,
             7, 7, 0, 1, ptr(drop[0]) if training else None, 4096, 0, 4096, st)
        h7 = torch.empty((B, hs, ws, npl * 4096), device=dev, dtype=tdtype)
        call("szn_conv_fwd", dt, ptr(h6), ptr(packed("fc7")), ptr(P["fc7.bias"].detach()), ptr(h7), B, hs, ws, 4096,
             4096, 1, 1, 0, 1, ptr(drop[1]) if training else None, 4096, 0, 4096, st)
        # score_fr and seenmask_score as ONE GEMM with N = D + 2 (padded to Dp)
        wf, ws_ = P["score_fr.weight"], P["seenmask_score.weight"]

        def build_head():
            cat = torch.cat([wf.detach(), ws_.detach()], 0).contiguous()
            bias = torch.zeros(Dp, device=dev, dtype=torch.float32)
            bias[:D] = P["score_fr.bias"].detach()
            bias[D:D + 2] = P["seenmask_score.bias"].detach()
            return _pack(dt, tdtype, cat, Dp), bias

        head_w, head_b = pw.get(("head", dt), (wf._version, ws_._version, P["score_fr.bias"]._version,
                                                P["seenmask_score.bias"]._version, wf.data_ptr(), ws_.data_ptr(),
                                                P["score_fr.bias"].data_ptr(), P["seenmask_score.bias"].data_ptr()),
                                  build_head)
        s17 = torch.empty((B, hs, ws, Dp), device=dev, dtype=torch.float32)
        call("szn_conv_fwd", dt, ptr(h7), ptr(head_w), ptr(head_b), ptr(s17), B, hs, ws, 4096, Dp, 1, 1, 0, 0,
             None, 0, 1, Dp, st)
        # upscore (x32 bilinear + crop 19) and the dense 2-channel seenmask_upscore
        up_w, sm_up_w = P["upscore.weight"], P["seenmask_upscore.weight"]
        diag = pw.get(("upscore_diag",), (up_w._version, up_w.data_ptr()), lambda: is_diag_bilinear(up_w.detach()))
        f = torch.empty((B, D, H, W), device=dev, dtype=torch.float32)
        if diag:
            call("szn_upsample32_crop_fwd", ptr(s17), ptr(f), B, D, H, W, hs, ws, Dp, 0, st)
        else:
            call("szn_deconv_small_fwd", ptr(s17), ptr(up_w.detach().contiguous()), ptr(f), B, D, D, H, W, hs, ws,
                 Dp, 0, st)
        s = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        call("szn_deconv_small_fwd", ptr(s17), ptr(sm_up_w.detach().contiguous()), ptr(s), B, 2, 2, H, W, hs, ws,
             Dp, D, st)

        ctx.module = module
        ctx.saved = dict(x=x, acts=acts, dims=dims, codes=codes, h6=h6, h7=h7, s17=s17, drop=drop, P=P, head_w=head_w,
                         geom=(B, H, W, D, Dp, hs, ws), diag=diag, dt=dt, tdtype=tdtype)
        if getattr(module, "fused_head", False) and diag:
            # experimental: also hand out the hs x ws score map, so that the fused head (utils._FusedHeadLoss) can send its
            # gradient straight back as d s17 without the (B, D, H, W) tensor ever being read
            # (a fresh view: the tensor kept in ctx.saved must not become an autograd output, or ctx -> s17 -> grad_fn ->
            # ctx would keep every activation of the step alive until the cycle collector runs)
            return f, s, s17.view(s17.shape)
        return f, s

    @staticmethod
    def backward(ctx, gf, gs, gs17=None):
        sv = ctx.saved
        module = ctx.module
        dt, tdtype = sv["dt"], sv["tdtype"]
        npl = planes(dt)
        st = _lib.stream()
        B, H, W, D, Dp, hs, ws = sv["geom"]
        P, acts, dims = sv["P"], sv["acts"], sv["dims"]
        dev = sv["x"].device
        pw = module._packed
        need = {n: g for n, g in zip(PARAM_ORDER, ctx.needs_input_grad[2:])}
        grads = _GradDict(module._grad_ready)
        for n in PARAM_ORDER:
            dict.__setitem__(grads, n, None)

        def packed_d(name, mode=0):
            w = P[name + ".weight"]
            return pw.get((name, "d", dt), (w._version, w.data_ptr()),
                          lambda: _pack_d(dt, tdtype, w.detach().contiguous(), mode))

        def zeros(shape, dtype=torch.float32):
            return torch.zeros(shape, device=dev, dtype=dtype)

        # ---------------- heads: d s17 ----------------
        if gf is None and gs is None and gs17 is None:
            return _finish(module, grads)
        ds17 = zeros((B, hs, ws, npl * Dp), tdtype)
        if gs17 is not None and gf is None:
            # fused head: d s17 arrives ready-made (fp32); store it in the trunk's gradient type (TF32-rounded / bf16).
            # Runs before the seen-mask head below, which overwrites its own two channels.
            call("szn_cast", dt, ptr(gs17.contiguous().float()), ptr(ds17), B * hs * ws, Dp, st)
        if gf is not None:
            gf = gf.contiguous()
            if sv["diag"]:
                call("szn_upsample32_crop_bwd", dt, ptr(gf), ptr(ds17), B, D, H, W, hs, ws, Dp, 0, st)
            else:
                call("szn_deconv_small_dgrad", dt, ptr(gf), ptr(P["upscore.weight"].detach().contiguous()), ptr(ds17),
                     B, D, D, H, W, hs, ws, Dp, 0, st)
            if need["upscore.weight"] and module.upscore_weight_grad:
                g = torch.empty_like(P["upscore.weight"])
                call("szn_deconv_small_wgrad", ptr(sv["s17"]), ptr(gf), ptr(g), B, D, D, H, W, hs, ws, Dp, 0, st)
                grads["upscore.weight"] = g
        if gs17 is not None and gf is not None:
            # the score was used both through the fused head and as a tensor: add the two contributions (rare)
            prev = ds17.float()
            if npl == 2:
                prev = prev[..., :Dp] + prev[..., Dp:]  # hi + lo planes
            both = (prev + gs17.float()).contiguous()
            call("szn_cast", dt, ptr(both), ptr(ds17), B * hs * ws, Dp, st)
        if gs is not None:
            gs = gs.contiguous()
            call("szn_deconv_small_dgrad", dt, ptr(gs), ptr(P["seenmask_upscore.weight"].detach().contiguous()),
                 ptr(ds17), B, 2, 2, H, W, hs, ws, Dp, D, st)
            if need["seenmask_upscore.weight"]:
                g = torch.empty_like(P["seenmask_upscore.weight"])
                call("szn_deconv_small_wgrad", ptr(sv["s17"]), ptr(gs), ptr(g), B, 2, 2, H, W, hs, ws, Dp, D, st)
                grads["seenmask_upscore.weight"] = g

        # which trunk layers still need a data gradient flowing into them
        trunk_need = [need[n + ".weight"] or need[n + ".bias"] for n in CONV_NAMES]
        first_needed = next((i for i, v in enumerate(trunk_need) if v), None)

        # ---------------- score heads: wgrad / bias / dgrad ----------------
        rows17 = B * hs * ws
        # a head whose output received no gradient (mode='fcn' / 'seenmask') gets None, as autograd gives the reference,
        # not zeros: an optimizer with momentum / weight decay would otherwise move parameters that were not used
        want = {"score_fr": gf is not None or gs17 is not None, "seenmask_score": gs is not None}
        if any(need[k + ".weight"] and want[k] for k in want):
            dwh = zeros((Dp, 4096))
            call("szn_conv_wgrad", dt, ptr(sv["h7"]), ptr(ds17), ptr(dwh), B, hs, ws, 4096, Dp, 1, 1, 0, Dp, st)
            if need["score_fr.weight"] and want["score_fr"]:
                grads["score_fr.weight"] = dwh[:D].reshape(D, 4096, 1, 1)
            if need["seenmask_score.weight"] and want["seenmask_score"]:
                grads["seenmask_score.weight"] = dwh[D:D + 2].reshape(2, 4096, 1, 1)
        if any(need[k + ".bias"] and want[k] for k in want):
            dbh = zeros((Dp,))
            call("szn_bias_grad", dt, ptr(ds17), ptr(dbh), rows17, Dp, Dp, st)
            if need["score_fr.bias"] and want["score_fr"]:
                grads["score_fr.bias"] = dbh[:D]
            if need["seenmask_score.bias"] and want["seenmask_score"]:
                grads["seenmask_score.bias"] = dbh[D:D + 2]
        if first_needed is None:
            return _finish(module, grads)

        # Bias gradients are column sums of a layer's dY.  The kernel that WRITES that dY (the next layer's dgrad epilogue
        # or the max-pool backward) accumulates them on the fly into `fused_db[layer]`, so dY is not read a second time.
        fused_db = {}

        def db_buffer(layer):
            if not need[layer + ".bias"] or _NO_FUSED_DB:
                return None
            fused_db[layer] = zeros((P[layer + ".bias"].shape[0],))
            return fused_db[layer]

        drop = sv["drop"]
        d7 = torch.empty((B, hs, ws, npl * 4096), device=dev, dtype=tdtype)
        wf, ws_ = P["score_fr.weight"], P["seenmask_score.weight"]
        head_wd = pw.get(("head", "d", dt), (wf._version, ws_._version, wf.data_ptr(), ws_.data_ptr()),
                         lambda: _pack_d(dt, tdtype, torch.cat([wf.detach(), ws_.detach()], 0).contiguous(), 0, Dp))
        call("szn_conv_dgrad", dt, ptr(ds17), ptr(head_wd), ptr(d7), B, hs, ws, 4096, Dp, 1, 1, 0,
             ptr(sv["h7"]), ptr(drop[1]) if drop is not None else None, 4096, Dp, ptr(db_buffer("fc7")), st)

        def conv_backward(name, x_act, dy, xh, xw, cin, cout, k, pad, want_dx, relu_ref, scale=None, producer=None):
            """wgrad + bias grad of one conv, then (optionally) its data gradient.  `producer`: the conv whose output
            (after ReLU, no pool in between) is this conv's input, i.e. whose dY the dgrad epilogue writes."""
            ho, wo = xh + 2 * pad - k + 1, xw + 2 * pad - k + 1
            if need[name + ".weight"]:
                dw = zeros((cout, k * k, cin))
                with _lib.nvtx_range(name + " wgrad"):
                    call("szn_conv_wgrad", dt, ptr(x_act), ptr(dy), ptr(dw), B, xh, xw, cin, cout, k, k, pad, cout, st)
                # dW lives as [Cout][R][S][Cin]; hand autograd the OIHW-shaped *view* of it (channels_last strides):
                # same values, no unpack pass over 135 M gradients
                grads[name + ".weight"] = dw.view(cout, k, k, cin).permute(0, 3, 1, 2)
            if need[name + ".bias"]:
                if name in fused_db:
                    grads[name + ".bias"] = fused_db.pop(name)
                else:
                    db = zeros((cout,))
                    call("szn_bias_grad", dt, ptr(dy), ptr(db), B * ho * wo, cout, cout, st)
                    grads[name + ".bias"] = db
            if not want_dx:
                return None
            dx = torch.empty((B, xh, xw, npl * cin), device=dev, dtype=tdtype)
            if k >= 5 and pad == 0 and relu_ref is None and scale is None:
                # fc6 (7x7 valid on 23x23): one GEMM against the (tap, ci)-major transposed weights gives per-tap
                # columns, which szn_col2im folds back; 98 full N tiles instead of 49 taps x a 512-wide N
                dcol = torch.empty((B, ho, wo, npl * k * k * cin), device=dev, dtype=tdtype)
                call("szn_conv_dgrad", dt, ptr(dy), ptr(packed_d(name, 1)), ptr(dcol), B, ho, wo, k * k * cin, cout, 1, 1,
                     0, None, None, 0, cout, None, st)
                call("szn_col2im", dt, ptr(dcol), ptr(dx), B, xh, xw, cin, k, k, st)
                return dx
            with _lib.nvtx_range(name + " dgrad"):
                call("szn_conv_dgrad", dt, ptr(dy), ptr(packed_d(name)), ptr(dx), B, xh, xw, cin, cout, k, k, pad,
                     ptr(relu_ref), ptr(scale), 4096 if scale is not None else 0, cout,
                     ptr(db_buffer(producer)) if producer is not None else None, st)
            return dx

        n_convs = len(CONV_NAMES)
        # fc7 (index n_convs-1), fc6 (n_convs-2)
        h5, w5, _ = dims["pool5"]
        d6 = conv_backward("fc7", sv["h6"], d7, hs, ws, 4096, 4096, 1, 0, first_needed <= n_convs - 2, sv["h6"],
                           drop[0] if drop is not None else None, producer="fc6")
        del d7
        if d6 is None:
            return _finish(module, grads)
        g = conv_backward("fc6", acts["pool5"], d6, h5, w5, 512, 4096, 7, 0, first_needed <= n_convs - 3, None)
        del d6
        # walk the trunk backwards
        rows = TRUNK
        i = len(rows) - 1
        while i >= 0 and g is not None:
            row = rows[i]
            if len(row) == 1:
                # g = d(pool out); route to the pre-pool activation (the previous conv's ReLU output)
                prev = rows[i - 1][0]
                ph, pw_, pc = dims[prev]
                dy = torch.empty((B, ph, pw_, npl * pc), device=dev, dtype=tdtype)
                vec = 8 if tdtype == torch.bfloat16 else 4
                fuse = 256 % (pc // vec) == 0
                if row[0] in sv["codes"]:
                    call("szn_pool_bwd_code", dt, ptr(sv["codes"][row[0]]), ptr(g), ptr(dy), B, ph, pw_, pc, 1,
                         ptr(db_buffer(prev)) if fuse else None, st)
                else:
                    call("szn_pool_bwd", dt, ptr(acts[prev]), ptr(g), ptr(dy), B, ph, pw_, pc, 1,
                         ptr(db_buffer(prev)) if fuse else None, st)
                g = dy
            else:
                name, cin, cout, k, pad = row
                ci = CONV_NAMES.index(name)
                if name == "conv1_1":
                    if need["conv1_1.weight"]:
                        dw = zeros((64, 3, 3, 3))
                        call("szn_conv1_1_wgrad", dt, ptr(sv["x"]), ptr(g), ptr(dw), B, H, W, 100, st)
                        grads["conv1_1.weight"] = dw
                    if need["conv1_1.bias"]:
                        if "conv1_1" in fused_db:
                            grads["conv1_1.bias"] = fused_db.pop("conv1_1")
                        else:
                            db = zeros((64,))
                            xh, xw, _ = dims["conv1_1"]
                            call("szn_bias_grad", dt, ptr(g), ptr(db), B * xh * xw, 64, 64, st)
                            grads["conv1_1.bias"] = db
                    g = None
                else:
                    prev = rows[i - 1][0]
                    xh, xw, _ = dims[prev]
                    prev_is_pool = len(rows[i - 1]) == 1
                    g = conv_backward(name, acts[prev], g, xh, xw, cin, cout, k, pad, first_needed < ci,
                                      None if prev_is_pool else acts[prev], producer=None if prev_is_pool else prev)
            i -= 1
        return _finish(module, grads)
