"""The algorithm of ``szn_pool_fwd_code / szn_pool_bwd_code`` (csrc/szn_simt.cu) restated in torch on the CPU and checked
against ATen's ``max_pool2d`` forward / backward (``models.py:47,54,63,72,81``: kernel 2, stride 2, ceil_mode=True):

* the four window loads are issued unconditionally with out-of-map positions CLAMPED onto the last row / column -- the
  clamped value repeats one seen earlier in scan order, so a strict ``>`` never lets it win and ``fmaxf`` is unchanged;
* code = winner (first maximum in scan order) | 4 if the maximum is positive (the ReLU gate of the producer conv);
* the backward pass needs dP and the code only.

The CUDA kernels themselves are compared bit for bit with this construction in tests/test_kernels_gpu.py::test_pool."""
import pytest
import torch
import torch.nn.functional as F


def pool_fwd_code(y):
    """y (B, C, H, W) -> pooled (B, C, Ho, Wo), code uint8 (B, C, Ho, Wo), with the kernel's clamped loads."""
    B, C, H, W = y.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    yo = torch.arange(Ho)[:, None]
    xo = torch.arange(Wo)[None, :]
    best = torch.full((B, C, Ho, Wo), float("-inf"))
    win = torch.zeros((B, C, Ho, Wo), dtype=torch.int64)
    for q in range(4):
        yy = (2 * yo + (q >> 1)).clamp(max=H - 1).expand(Ho, Wo)
        xx = (2 * xo + (q & 1)).clamp(max=W - 1).expand(Ho, Wo)
        v = y[:, :, yy, xx]
        take = v > best                       # strict: the first maximum in scan order keeps the gradient
        best = torch.where(take, v, best)
        win = torch.where(take, torch.full_like(win, q), win)
    return best, (win + 4 * (best > 0)).to(torch.uint8)


def pool_bwd_code(code, dp, H, W, relu_gate):
    B, C, Ho, Wo = dp.shape
    dy = torch.zeros((B, C, H, W))
    for q in range(4):
        ys = torch.arange(Ho) * 2 + (q >> 1)
        xs = torch.arange(Wo) * 2 + (q & 1)
        oky, okx = ys < H, xs < W
        sel = ((code & 3) == q) & (((code & 4) != 0) | (not relu_gate))
        g = torch.where(sel, dp, torch.zeros_like(dp))[:, :, oky][:, :, :, okx]
        dy[:, :, ys[oky][:, None], xs[okx][None, :]] = g
    return dy


@pytest.mark.parametrize("hw", [(1, 1), (2, 3), (7, 9), (8, 8), (45, 23), (5, 1)])
def test_routing_codes_reproduce_aten(hw):
    H, W = hw
    g = torch.Generator().manual_seed(100 + H * 31 + W)
    y = F.relu(torch.randn(2, 5, H, W, generator=g))          # a ReLU output: many exact zeros (ties at 0)
    if H > 1 and W > 1:
        y[:, :, 0:2, 0:2] = y[:, :, 0:1, 0:1]                   # exact positive ties: the first position must win
    y.requires_grad_(True)
    p = F.max_pool2d(y, 2, stride=2, ceil_mode=True)
    pooled, code = pool_fwd_code(y.detach())
    assert torch.equal(pooled, p.detach())
    assert int(code.max()) <= 7
    dp = torch.randn(p.shape, generator=g)
    (dy,) = torch.autograd.grad(p, y, dp)
    assert torch.equal(pool_bwd_code(code, dp, H, W, relu_gate=False), dy)
    # with the gate: positions whose activation is 0 (ReLU inactive) get no gradient -- what szn_pool_bwd computed from y
    assert torch.equal(pool_bwd_code(code, dp, H, W, relu_gate=True), dy * (y.detach() > 0))


def test_a_clamped_duplicate_never_wins():
    # odd map: the last window holds ONE real value; its three clamped copies must not take the gradient
    y = torch.tensor([[[[1.0, 2.0, 3.0], [4.0, 5.0, 6.0], [7.0, 8.0, 9.0]]]])
    pooled, code = pool_fwd_code(y)
    assert torch.equal(pooled, torch.tensor([[[[5.0, 6.0], [8.0, 9.0]]]]))
    assert code.tolist() == [[[[3 + 4, 2 + 4], [1 + 4, 0 + 4]]]]   # corner window: winner 0 (the only real position)
