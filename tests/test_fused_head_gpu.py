"""Fused head (FCN32s(fused_head=True), szn_head_fused_*): loss, labels and d s17 from the 17x17 score map against the
ordinary path that materialises the (B, D, H, W) score (the parity reference; SURVEY 7 "commuting the head").  Green on a
B200 since the first GPU call of round 2."""
import numpy as np
import pytest
import torch

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("B,D,C,H,W", [(2, 20, 21, 40, 56), (1, 300, 59, 96, 64), (3, 5, 7, 33, 17), (1, 8, 3, 5, 3)])
def test_fused_kernels_equal_materialised_head(B, D, C, H, W):
    from zeroshotsemanticsegmentation_b200 import _lib, utils as U
    def score_map(n):  # conv1_1 pad 100 (+198), five ceil-mode 2x2 pools, fc6 7x7 valid (-6)
        n += 198
        for _ in range(5):
            n = (n + 1) // 2
        return n - 6
    hs, ws = score_map(H), score_map(W)
    Dp = (D + 2 + 63) // 64 * 64
    g = torch.Generator().manual_seed(B * 100 + D)
    s17 = torch.randn(B, hs, ws, Dp, generator=g).to(DEV)
    _, lab, table = O.synth_batch(B, H, W, C, D, seed=D, block=4, ignore_frac=0.1)
    lab, table = lab.to(DEV), table.to(DEV)
    # ordinary path: upsample kernel -> loss / labels kernels on the materialised score
    f = torch.empty(B, D, H, W, device=DEV)
    _lib.call("szn_upsample32_crop_fwd", s17.data_ptr(), f.data_ptr(), B, D, H, W, hs, ws, Dp, 0, st())
    f.requires_grad_(True)
    loss_ref = U.cosine_loss(f, lab, table=table)
    (gf,) = torch.autograd.grad(loss_ref, f)
    ds_ref = torch.zeros(B, hs, ws, Dp, device=DEV)
    _lib.call("szn_upsample32_crop_bwd", 0, gf.data_ptr(), ds_ref.data_ptr(), B, D, H, W, hs, ws, Dp, 0, st())
    lbl_ref = U.infer_lbl_device(f.detach(), table)
    # fused path
    s = s17.clone().requires_grad_(True)
    loss = U._FusedHeadLoss.apply(s, lab, table, D, (H, W), None)
    (ds,) = torch.autograd.grad(loss, s)
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    err = float((ds[..., :D] - ds_ref[..., :D]).norm() / ds_ref[..., :D].norm())
    print("d s17 rel-L2 error %.3e" % err)
    assert err < 1e-3  # ds_ref is stored rounded to TF32
    assert float(ds[..., D:].abs().max()) == 0.0
    work = U._fused_workspace(s17, C)
    lbl = torch.empty(B, H, W, dtype=torch.int64, device=DEV)
    _lib.call("szn_head_fused_fwd", 0, s17.data_ptr(), Dp, 0, None, table.data_ptr(), B, D, H, W, hs, ws, C, work.data_ptr(),
              None, None, lbl.data_ptr(), st())
    torch.cuda.synchronize()
    assert float((lbl != lbl_ref).float().mean()) < 1e-3  # another summation order: near-ties only
    # the MSE variant (utils.py:50-73) from the same quantities
    f2 = f.detach().clone().requires_grad_(True)
    mse_ref = U.mse_loss(f2, lab, table=table)
    (gf2,) = torch.autograd.grad(mse_ref, f2)
    ds_ref2 = torch.zeros(B, hs, ws, Dp, device=DEV)
    _lib.call("szn_upsample32_crop_bwd", 0, gf2.data_ptr(), ds_ref2.data_ptr(), B, D, H, W, hs, ws, Dp, 0, st())
    s2 = s17.clone().requires_grad_(True)
    mse = U._FusedHeadLoss.apply(s2, lab, table, D, (H, W), None, 1)
    (ds2,) = torch.autograd.grad(mse, s2)
    assert abs(mse.item() - mse_ref.item()) < 1e-4 * max(1.0, abs(mse_ref.item()))
    assert float((ds2[..., :D] - ds_ref2[..., :D]).norm() / ds_ref2[..., :D].norm()) < 1e-3


def test_model_with_fused_head_equals_default_path():
    import zeroshotsemanticsegmentation_b200 as szn
    U = szn.utils
    D, C, H, W, B = 20, 21, 40, 56, 2
    params = O.init_params(D, seed=31)
    x, lab, table = O.synth_batch(B, H, W, C, D, seed=31, block=8)
    x, lab, table = x.to(DEV), lab.to(DEV), table.to(DEV)
    res = {}
    for fused in (False, True):
        m = szn.FCN32s(D, fused_head=fused)
        m.load_state_dict(params)
        m = m.to(DEV).eval()
        f = m(x, mode="fcn")
        assert (getattr(f, "_szn_head", None) is not None) == fused
        loss = U.cosine_loss(f, lab, table=table)
        loss.backward()
        lbl = U.infer_lbl_device(f, table)
        res[fused] = (f.detach(), loss.item(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, lbl)
    assert torch.equal(res[True][0], res[False][0])
    assert abs(res[True][1] - res[False][1]) < 1e-5
    assert set(res[True][2]) == set(res[False][2])
    for n, gref in res[False][2].items():
        e = float((res[True][2][n] - gref).norm() / gref.norm().clamp_min(1e-30))
        assert e < (1e-2 if n.startswith(("score_fr", "fc7")) else 0.3), (n, e)
    assert float((res[True][3] != res[False][3]).float().mean()) < 1e-3
    # a modified score falls back to the ordinary path
    m = szn.FCN32s(D, fused_head=True)
    m.load_state_dict(params)
    m = m.to(DEV).eval()
    f = m(x, mode="fcn")
    assert U._fused_handle(f) is not None and U._fused_handle(f.detach()) is None and U._fused_handle(f * 1.0) is None
    # the loss pass leaves the labels of (this score, this table) behind; infer_lbl picks them up, another table does not
    head = U._fused_handle(f)
    assert getattr(head, "labels_cache", None) is None
    fresh = U.infer_lbl_device(f, table)            # label-only launch
    U.cosine_loss(f, lab, table=table)
    assert head.labels_cache is not None
    cached = U.infer_lbl_device(f, table)
    assert cached is head.labels_cache[3] and torch.equal(cached, fresh)
    other = table.flip(0).contiguous()
    assert not torch.equal(U.infer_lbl_device(f, other), cached)
