"""Per-kernel parity tests (GPU): every C-ABI op of libszn.so against the CPU fp32 oracle ops.

Inputs are pre-rounded to the kernel's storage type (TF32 / bf16), so the tensor-core products are exact
and the only difference to the fp32 CPU result is accumulation order: tolerances are tight (1e-4 .. 1e-3)
and any layout / descriptor / indexing bug shows up as O(1) error.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def lib():
    from zeroshotsemanticsegmentation_b200 import _lib
    return _lib


PRECS = ["tf32", "bf16", "fp32"]  # "fp32" = the split (hi, lo) bf16-plane format, 3 x bf16 products
# kernels new in this build (csrc/szn_internal.h SZN_NEW_KERNELS_DEFAULT) are tested when they are the default or on request
import os
NEW_KERNELS = os.environ.get("SZN_TEST_NEW", "1") == "1"  # SZN_TEST_NEW=0 skips the variants that are not the default
needs_new = pytest.mark.skipif(not NEW_KERNELS, reason="new-kernel variants: SZN_TEST_NEW=1")


def round_to(t, prec):
    """Round to what the storage format represents exactly."""
    if prec == "bf16":
        return t.bfloat16().float()
    if prec == "fp32":
        hi = t.bfloat16().float()
        return hi + (t - hi).bfloat16().float()
    i = t.contiguous().view(torch.int32)
    i = ((i + 0xFFF + ((i >> 13) & 1)) >> 13) << 13
    return i.view(torch.float32)


def tdtype(prec):
    return torch.float32 if prec == "tf32" else torch.bfloat16


def dcode(prec):
    return {"tf32": 0, "bf16": 1, "fp32": 2}[prec]


def planes(prec):
    return 2 if prec == "fp32" else 1


def to_store(t, prec):
    """fp32 [..., C] -> device tensor in the storage format ([..., 2C] bf16 = hi | lo for the split format)."""
    if prec == "fp32":
        hi = t.bfloat16()
        lo = (t - hi.float()).bfloat16()
        return torch.cat([hi, lo], -1).contiguous().to(DEV)
    return t.contiguous().to(DEV).to(tdtype(prec))


def from_store(t, prec):
    t = t.float().cpu()
    if prec == "fp32":
        C = t.shape[-1] // 2
        return t[..., :C] + t[..., C:]
    return t


def act_buf(prec, *shape, fill=float("nan")):
    """Uninitialised (NaN-filled) activation buffer [..., C] in the storage format."""
    shape = list(shape)
    shape[-1] *= planes(prec)
    return torch.full(shape, fill, device=DEV, dtype=tdtype(prec))


def nhwc(t, prec):  # NCHW fp32 cpu -> NHWC device tensor of the storage type
    return to_store(t.permute(0, 2, 3, 1), prec)


def from_nhwc(t, prec="tf32"):
    return from_store(t, prec).permute(0, 3, 1, 2).contiguous()


def ohwi(w, prec, o_pad=None):
    O_, I, R, S = w.shape
    out = w.permute(0, 2, 3, 1).reshape(O_, R * S, I)
    if o_pad and o_pad > O_:
        out = torch.cat([out, torch.zeros(o_pad - O_, R * S, I)], 0)
    if prec == "fp32":  # two planes, hi then lo
        hi = out.bfloat16()
        return torch.cat([hi, (out - hi.float()).bfloat16()], 0).contiguous().to(DEV)
    return out.contiguous().to(DEV).to(tdtype(prec))


TOL_STORE = {"tf32": 1e-3, "bf16": 1e-2, "fp32": 4e-5}  # outputs rounded to the storage type


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def st():
    return torch.cuda.current_stream().cuda_stream


_KEEP = []


def dp(t):
    """Device pointer of a tensor that is kept alive until the test ends (a temporary would be freed -- and its
    block reused by the next allocation -- before the kernel that reads it has even been enqueued)."""
    _KEEP.append(t)
    return t.data_ptr()


@pytest.fixture(autouse=True)
def _release_after_test():
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


CONV_CASES = [
    # B, H, W, Cin, Cout, k, pad
    (2, 20, 37, 64, 64, 3, 1),
    (1, 9, 9, 128, 256, 3, 1),
    (1, 45, 45, 64, 128, 3, 1),
    (2, 9, 10, 64, 128, 7, 0),
    (2, 5, 7, 256, 320, 1, 0),
    (1, 23, 23, 128, 512, 3, 1),
    (2, 33, 50, 64, 64, 3, 1),    # 3x3 halo-patch mode with the whole filter resident in shared memory (conv1_2-like)
    (1, 48, 40, 256, 128, 3, 1),  # halo-patch mode, streamed weights, several tiles per CTA
]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd(prec, case):
    B, H, W, Cin, Cout, k, pad = case
    g = torch.Generator().manual_seed(1)
    x = round_to(torch.randn(B, Cin, H, W, generator=g), prec)
    w = round_to(torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5, prec)
    b = torch.randn(Cout, generator=g)
    scale = (torch.rand(B, Cout, generator=g) < 0.5).float() * 2
    ref = F.relu(F.conv2d(x, w, b, padding=pad)) * scale[:, :, None, None]
    Ho, Wo = ref.shape[2:]
    y = act_buf(prec, B, Ho, Wo, Cout)
    L = lib()
    L.call("szn_conv_fwd", dcode(prec), dp(nhwc(x, prec)), dp(ohwi(w, prec)), dp(b.to(DEV)),
           dp(y), B, H, W, Cin, Cout, k, k, pad, 1, dp(scale.to(DEV)), Cout, 0, Cout, st())
    torch.cuda.synchronize()
    got = from_nhwc(y, prec)
    tol = TOL_STORE[prec]  # the kernel rounds its output to the storage type
    e = relerr(got, ref)
    print("conv_fwd", prec, case, "relerr", e)
    assert e < tol


@pytest.mark.parametrize("prec", PRECS)
def test_conv_fwd_fp32_out_no_relu(prec):
    B, H, W, Cin, Cout = 2, 5, 7, 128, 64
    g = torch.Generator().manual_seed(2)
    x = round_to(torch.randn(B, Cin, H, W, generator=g), prec)
    w = round_to(torch.randn(Cout - 10, Cin, 1, 1, generator=g) / Cin ** 0.5, prec)
    b = torch.cat([torch.randn(Cout - 10, generator=g), torch.zeros(10)])
    ref = F.conv2d(x, w, b[:Cout - 10])
    y = torch.full((B, H, W, Cout), float("nan"), device=DEV, dtype=torch.float32)
    L = lib()
    L.call("szn_conv_fwd", dcode(prec), dp(nhwc(x, prec)), dp(ohwi(w, prec, Cout)), dp(b.to(DEV)),
           dp(y), B, H, W, Cin, Cout, 1, 1, 0, 0, None, 0, 1, Cout, st())
    torch.cuda.synchronize()
    got = from_nhwc(y)
    assert relerr(got[:, :Cout - 10], ref) < 1e-5
    assert (got[:, Cout - 10:] == 0).all()


def pack_dgrad(w, prec, mode=0):
    """Transposed data-gradient weights through the library's own pack kernel (what engine.py does)."""
    O_, I, R, S = w.shape
    out = torch.empty((planes(prec) * I, R * S, O_), device=DEV, dtype=tdtype(prec))
    lib().call("szn_pack_weight_dgrad", dcode(prec), dp(w.contiguous().to(DEV)), dp(out), O_, I, R, S, O_, mode, st())
    return out


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad(prec, case):
    B, H, W, Cin, Cout, k, pad = case
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, Cin, H, W, generator=g).requires_grad_(True)
    w = round_to(torch.randn(Cout, Cin, k, k, generator=g) / (Cout * k * k) ** 0.5, prec)
    y = F.conv2d(x, w, padding=pad)
    dy = round_to(torch.randn(y.shape, generator=g), prec)
    ref_act = torch.randn(B, Cin, H, W, generator=g)  # the layer input whose sign gates the gradient
    scale = (torch.rand(B, Cin, generator=g) < 0.5).float() * 2
    (dx,) = torch.autograd.grad(y, x, dy)
    ref = dx * scale[:, :, None, None] * (ref_act > 0)
    out = act_buf(prec, B, H, W, Cin)
    colsum = torch.zeros(Cin, device=DEV)
    L = lib()
    L.call("szn_conv_dgrad", dcode(prec), dp(nhwc(dy, prec)), dp(pack_dgrad(w, prec)), dp(out), B, H, W,
           Cin, Cout, k, k, pad, dp(nhwc(ref_act, prec)), dp(scale.to(DEV)), Cin, Cout, dp(colsum), st())
    torch.cuda.synchronize()
    got = from_nhwc(out, prec)
    e = relerr(got, ref)
    print("conv_dgrad", prec, case, "relerr", e)
    assert e < TOL_STORE[prec]
    # fused bias gradient of the producer layer: column sums of exactly the values that were stored
    want = got.sum(dim=(0, 2, 3))
    assert relerr(colsum.cpu(), want) < 1e-5


@pytest.mark.parametrize("prec", PRECS)
def test_conv_dgrad_col2im(prec):
    """fc6-style data gradient: one GEMM against the (tap, ci)-major transposed weights + szn_col2im."""
    B, H, W, Cin, Cout, k = 2, 11, 12, 64, 128, 7
    g = torch.Generator().manual_seed(13)
    x = torch.randn(B, Cin, H, W, generator=g).requires_grad_(True)
    w = round_to(torch.randn(Cout, Cin, k, k, generator=g) / (Cout * k * k) ** 0.5, prec)
    y = F.conv2d(x, w)
    dy = round_to(torch.randn(y.shape, generator=g), prec)
    (ref,) = torch.autograd.grad(y, x, dy)
    Ho, Wo = y.shape[2:]
    L = lib()
    dcol = act_buf(prec, B, Ho, Wo, k * k * Cin)
    L.call("szn_conv_dgrad", dcode(prec), dp(nhwc(dy, prec)), dp(pack_dgrad(w, prec, 1)), dp(dcol), B, Ho, Wo,
           k * k * Cin, Cout, 1, 1, 0, None, None, 0, Cout, None, st())
    out = act_buf(prec, B, H, W, Cin)
    L.call("szn_col2im", dcode(prec), dp(dcol), dp(out), B, H, W, Cin, k, k, st())
    torch.cuda.synchronize()
    e = relerr(from_nhwc(out, prec), ref)
    print("conv_dgrad_col2im", prec, "relerr", e)
    assert e < {"tf32": 1e-3, "bf16": 2e-2, "fp32": 1e-4}[prec]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_wgrad(prec, case):
    B, H, W, Cin, Cout, k, pad = case
    g = torch.Generator().manual_seed(4)
    x = round_to(torch.randn(B, Cin, H, W, generator=g), prec)
    w = torch.zeros(Cout, Cin, k, k, requires_grad=True)
    y = F.conv2d(x, w, padding=pad)
    dy = round_to(torch.randn(y.shape, generator=g), prec)
    (dw,) = torch.autograd.grad(y, w, dy)
    ref = dw.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)
    out = torch.zeros((Cout, k * k * Cin), device=DEV, dtype=torch.float32)
    L = lib()
    L.call("szn_conv_wgrad", dcode(prec), dp(nhwc(x, prec)), dp(nhwc(dy, prec)), dp(out), B, H, W,
           Cin, Cout, k, k, pad, Cout, st())
    torch.cuda.synchronize()
    e = relerr(out.cpu(), ref)
    print("conv_wgrad", prec, case, "relerr", e)
    assert e < 1e-4


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("ctas", [pytest.param("2", marks=needs_new), "1"])
def test_conv1_1(prec, ctas, monkeypatch):
    # the tensor-core forward with two CTAs per SM (half-tile staging) and one (full-tile staging); the weight gradient in
    # the re-blocked form (with "2") and the older one
    monkeypatch.setenv("SZN_CONV1_1_TC_CTAS", ctas)
    monkeypatch.setenv("SZN_CONV1_1_WGRAD_V2", "1" if ctas == "2" else "0")
    B, H, W = 2, 13, 21
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, H, W, generator=g) * 50
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.1).requires_grad_(True)
    b = torch.randn(64, generator=g)
    y = F.relu(F.conv2d(x, w, b, padding=100))
    Ho, Wo = y.shape[2:]
    out = act_buf(prec, B, Ho, Wo, 64)
    L = lib()
    L.call("szn_conv1_1_fwd", dcode(prec), dp(x.to(DEV)), dp(w.detach().to(DEV)), dp(b.to(DEV)),
           dp(out), B, H, W, 100, st())
    torch.cuda.synchronize()
    assert relerr(from_nhwc(out, prec), y.detach()) < TOL_STORE[prec]
    dy = round_to(torch.randn(y.shape, generator=g), prec)
    pre = F.conv2d(x, w, padding=100)
    (dw,) = torch.autograd.grad(pre, w, dy)
    gw = torch.zeros((64, 3, 3, 3), device=DEV)
    L.call("szn_conv1_1_wgrad", dcode(prec), dp(x.to(DEV)), dp(nhwc(dy, prec)), dp(gw), B, H, W, 100,
           st())
    torch.cuda.synchronize()
    assert relerr(gw.cpu(), dw) < 1e-4


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("v2", [pytest.param(True, marks=needs_new), False])
@pytest.mark.parametrize("case", [(5, 130, 100), (9, 62, 100), (6, 63, 100), (8, 70, 1), (3, 200, 2)])
def test_conv1_1_wgrad_row_segments(prec, case, v2, monkeypatch):
    """Rows of the contributing window that span several 64-pixel segments (full, exactly full, 1- and 4-pixel tails) and
    other paddings, through the re-blocked kernel and the older (co, ci, filter row) blocking (SZN_CONV1_1_WGRAD_V2, read
    per call)."""
    H, W, pad = case
    B = 2
    g = torch.Generator().manual_seed(50 + H)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.zeros(64, 3, 3, 3, requires_grad=True)
    pre = F.conv2d(x, w, padding=pad)
    dy = round_to(torch.randn(pre.shape, generator=g), prec)
    (dw,) = torch.autograd.grad(pre, w, dy)
    monkeypatch.setenv("SZN_CONV1_1_WGRAD_V2", "1" if v2 else "0")
    gw = torch.zeros((64, 3, 3, 3), device=DEV)
    lib().call("szn_conv1_1_wgrad", dcode(prec), dp(x.to(DEV)), dp(nhwc(dy, prec)), dp(gw), B, H, W, pad, st())
    torch.cuda.synchronize()
    assert relerr(gw.cpu(), dw) < 1e-4


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("hw", [(7, 9), (8, 8), (45, 23)])
def test_pool(prec, hw):
    B, C = 2, 64
    H, W = hw
    g = torch.Generator().manual_seed(6)
    # ReLU output with many exact zeros and exact ties inside windows
    y = round_to(F.relu(torch.randn(B, C, H, W, generator=g)), prec)
    y[:, :, 0:2, 0:2] = y[:, :, 0:1, 0:1]  # exact non-zero ties: the first position must win
    y.requires_grad_(True)
    p = F.max_pool2d(y, 2, stride=2, ceil_mode=True)
    out = act_buf(prec, B, p.shape[2], p.shape[3], C)
    L = lib()
    yd = nhwc(y.detach(), prec)
    L.call("szn_pool_fwd", dcode(prec), dp(yd), dp(out), B, H, W, C, st())
    torch.cuda.synchronize()
    assert torch.equal(from_nhwc(out, prec), p.detach())
    dpool = round_to(torch.randn(p.shape, generator=g), prec)
    (dy,) = torch.autograd.grad(p, y, dpool)
    ref = dy * (y.detach() > 0)
    dyo = act_buf(prec, B, H, W, C)
    csum = torch.zeros(C, device=DEV)
    L.call("szn_pool_bwd", dcode(prec), dp(yd), dp(nhwc(dpool, prec)), dp(dyo), B, H, W, C, 1, dp(csum), st())
    torch.cuda.synchronize()
    assert torch.equal(from_nhwc(dyo, prec), ref)
    assert relerr(csum.cpu(), ref.sum(dim=(0, 2, 3))) < 1e-5  # fused bias gradient of the conv in front of the pool
    if not NEW_KERNELS:
        return
    # the routing-code pair: one byte per pooled element instead of re-reading y (bit-identical results)
    out2 = act_buf(prec, B, p.shape[2], p.shape[3], C)
    code = torch.full((B, p.shape[2], p.shape[3], C), 255, device=DEV, dtype=torch.uint8)
    L.call("szn_pool_fwd_code", dcode(prec), dp(yd), dp(out2), dp(code), B, H, W, C, st())
    torch.cuda.synchronize()
    assert torch.equal(out2, out)
    yc = y.detach()
    # reference codes: first maximum in scan order (ATen's choice) | 4 where it is positive
    win = torch.zeros(p.shape, dtype=torch.int64)
    best = torch.full(p.shape, float("-inf"))
    for q in range(4):
        sub = yc[:, :, (q >> 1)::2, (q & 1)::2]
        pad_ = torch.full(p.shape, float("-inf"))
        pad_[:, :, :sub.shape[2], :sub.shape[3]] = sub
        take = pad_ > best
        win[take], best[take] = q, pad_[take]
    ref_code = (win + 4 * (best > 0)).permute(0, 2, 3, 1).to(torch.uint8)
    assert torch.equal(code.cpu(), ref_code)
    dyo2 = act_buf(prec, B, H, W, C)
    csum2 = torch.zeros(C, device=DEV)
    L.call("szn_pool_bwd_code", dcode(prec), dp(code), dp(nhwc(dpool, prec)), dp(dyo2), B, H, W, C, 1, dp(csum2), st())
    torch.cuda.synchronize()
    assert torch.equal(dyo2, dyo)
    assert torch.equal(from_nhwc(dyo2, prec), ref)
    assert relerr(csum2.cpu(), ref.sum(dim=(0, 2, 3))) < 1e-5
    # without the ReLU gate the routed gradient is ATen's max_pool2d backward itself
    L.call("szn_pool_bwd_code", dcode(prec), dp(code), dp(nhwc(dpool, prec)), dp(dyo2), B, H, W, C, 0, None, st())
    torch.cuda.synchronize()
    assert torch.equal(from_nhwc(dyo2, prec), dy)


@pytest.mark.parametrize("prec", PRECS)
def test_bias_grad_pack_unpack(prec):
    g = torch.Generator().manual_seed(7)
    rows, C, ld = 1000, 320, 320
    dy = round_to(torch.randn(rows, ld, generator=g), prec)
    db = torch.zeros(C, device=DEV)
    L = lib()
    L.call("szn_bias_grad", dcode(prec), dp(to_store(dy, prec)), dp(db), rows, C, ld, st())
    torch.cuda.synchronize()
    assert relerr(db.cpu(), dy.sum(0)) < 1e-5
    w = torch.randn(5, 64, 3, 3, generator=g)
    out = torch.empty((planes(prec) * 8, 9, 64), device=DEV, dtype=tdtype(prec))
    L.call("szn_pack_weight", dcode(prec), dp(w.to(DEV)), dp(out), 5, 64, 3, 3, 8, st())
    torch.cuda.synchronize()
    ref = round_to(w, prec).permute(0, 2, 3, 1).reshape(5, 9, 64)
    got = out.float().cpu()
    if prec == "fp32":
        got = got[:8] + got[8:]  # hi plane + lo plane
    assert torch.equal(got[:5], ref) and (got[5:] == 0).all()
    dw = torch.randn(5, 9, 64, generator=g)
    gg = torch.empty((5, 64, 3, 3), device=DEV)
    L.call("szn_unpack_wgrad", dp(dw.to(DEV)), dp(gg), 5, 64, 3, 3, st())
    torch.cuda.synchronize()
    assert torch.equal(gg.cpu(), dw.reshape(5, 3, 3, 64).permute(0, 3, 1, 2))


# native-size PASCAL images (train.py:82-84 feeds them unpadded at batch size 1): W % 4 != 0 and W > 256 takes the
# multi-slot VEC=1 backward kernel
@pytest.mark.parametrize("HW", [(37, 53), (64, 96), (500, 375), (375, 500), (333, 486), (40, 990)])
def test_upsample_and_small_deconv(HW):
    H, W = HW
    B, D = (2, 20) if H * W < 20000 else (1, 5)
    hs, ws = (H + 198 + 31) // 32 - 6, (W + 198 + 31) // 32 - 6
    # trunk geometry: 5 ceil-mode pools then 7x7 valid
    def geom(n):
        n = n + 198
        for _ in range(5):
            n = (n + 1) // 2
        return n - 6
    hs, ws = geom(H), geom(W)
    ld = 32
    g = torch.Generator().manual_seed(8)
    s = torch.randn(B, hs, ws, ld, generator=g)
    s_nchw = s.permute(0, 3, 1, 2).contiguous()
    wdiag = O.upsampling_weight(D, D)
    sref = s_nchw[:, :D].clone().requires_grad_(True)
    full = F.conv_transpose2d(sref, wdiag, stride=32)[:, :, 19:19 + H, 19:19 + W]
    L = lib()
    out = torch.empty((B, D, H, W), device=DEV)
    L.call("szn_upsample32_crop_fwd", dp(s.to(DEV)), dp(out), B, D, H, W, hs, ws, ld, 0, st())
    torch.cuda.synchronize()
    assert relerr(out.cpu(), full.detach()) < 1e-5
    gout = torch.randn(B, D, H, W, generator=g)
    (ds_ref,) = torch.autograd.grad(full, sref, gout)
    ds = torch.zeros((B, hs, ws, ld), device=DEV)
    L.call("szn_upsample32_crop_bwd", 0, dp(gout.to(DEV)), dp(ds), B, D, H, W, hs, ws, ld, 0, st())
    torch.cuda.synchronize()
    got = ds.cpu().permute(0, 3, 1, 2)[:, :D]
    assert relerr(got, round_to(ds_ref, "tf32")) < 1e-3
    # dense 2x2 head at channel offset D
    wd = torch.randn(2, 2, 64, 64, generator=g).requires_grad_(True)
    s2 = s_nchw[:, D:D + 2].clone().requires_grad_(True)
    y2 = F.conv_transpose2d(s2, wd, stride=32)[:, :, 19:19 + H, 19:19 + W]
    o2 = torch.empty((B, 2, H, W), device=DEV)
    L.call("szn_deconv_small_fwd", dp(s.to(DEV)), dp(wd.detach().to(DEV)), dp(o2), B, 2, 2, H, W,
           hs, ws, ld, D, st())
    torch.cuda.synchronize()
    assert relerr(o2.cpu(), y2.detach()) < 1e-5
    g2 = torch.randn(B, 2, H, W, generator=g)
    ds2_ref, dwd_ref = torch.autograd.grad(y2, (s2, wd), g2)
    ds2 = torch.zeros((B, hs, ws, ld), device=DEV)
    L.call("szn_deconv_small_dgrad", 0, dp(g2.to(DEV)), dp(wd.detach().to(DEV)), dp(ds2), B, 2, 2, H,
           W, hs, ws, ld, D, st())
    dwd = torch.empty((2, 2, 64, 64), device=DEV)
    L.call("szn_deconv_small_wgrad", dp(s.to(DEV)), dp(g2.to(DEV)), dp(dwd), B, 2, 2, H, W, hs, ws,
           ld, D, st())
    torch.cuda.synchronize()
    assert relerr(ds2.cpu().permute(0, 3, 1, 2)[:, D:D + 2], round_to(ds2_ref, "tf32")) < 1e-3
    assert relerr(dwd.cpu(), dwd_ref) < 1e-4
