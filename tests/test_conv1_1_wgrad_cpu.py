"""Index-for-index restatement of ``conv1_1_wgrad_v2_kernel`` (csrc/szn_simt.cu) in numpy: the contributing window of
output pixels, its rows cut into 64-pixel segments, the zero-padded staging of a dY tile and of the three input rows with
their halo, and the quad loop that visits ceil(npx / 4) pixel quads -- against autograd's weight gradient of
``Conv2d(3, 64, 3, padding=pad)`` (``models.py:43``).  The CUDA kernel is compared with the same autograd result in
tests/test_kernels_gpu.py::test_conv1_1_wgrad_row_segments; this pins the bookkeeping on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

SEG, XROW = 64, 68


def wgrad_v2(x, dy, pad):
    """x (B,3,H,W), dy (B,64,Ho,Wo) -> dw (64,3,3,3), walking work items like the kernel (host code of szn_conv1_1_wgrad)."""
    B, _, H, W = x.shape
    Ho, Wo = H + 2 * pad - 2, W + 2 * pad - 2
    ylo = max(pad - 2, 0)
    xlo = ylo
    yhi, xhi = min(pad + H - 1, Ho - 1), min(pad + W - 1, Wo - 1)
    wy, wx = yhi - ylo + 1, xhi - xlo + 1            # output pixels whose 3x3 window touches the image
    nseg = (wx + SEG - 1) // SEG
    dyn = dy.permute(0, 2, 3, 1).numpy()             # NHWC, as the kernel reads it
    xn = x.numpy()
    acc = np.zeros((64, 3, 3, 3), dtype=np.float64)  # [co][ci][r][t]
    for it in range(B * wy * nseg):
        seg, yo, b = it % nseg, ylo + (it // nseg) % wy, it // (nseg * wy)
        xo0 = xlo + seg * SEG
        npx = min(Wo - xo0, SEG)
        sdy = np.zeros((SEG, 64))
        sdy[:npx] = dyn[b, yo, xo0:xo0 + npx]        # pixels past the row end count as 0
        xs = np.zeros((9, XROW))
        for row in range(9):
            ci, r = divmod(row, 3)
            yi = yo + r - pad
            for col in range(XROW):
                xi = xo0 - pad + col
                if 0 <= yi < H and 0 <= xi < W:
                    xs[row, col] = xn[b, ci, yi, xi]
        nq = (npx + 3) // 4
        for k in range(nq):
            for j in range(4):
                px = 4 * k + j
                for t in range(3):
                    # acc[co][ci][r][t] += dy[px][co] * xs[ci*3 + r][px + t]
                    acc[:, :, :, t] += sdy[px][:, None, None] * xs[:, px + t].reshape(3, 3)[None]
    return torch.from_numpy(acc).float()


@pytest.mark.parametrize("case", [(2, 5, 130, 100), (1, 9, 62, 100), (1, 6, 63, 100), (2, 8, 70, 1), (1, 3, 200, 2),
                                  (1, 4, 5, 0 + 2)])
def test_restated_kernel_equals_autograd(case):
    B, H, W, pad = case
    g = torch.Generator().manual_seed(7 * H + W)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.zeros(64, 3, 3, 3, requires_grad=True)
    pre = F.conv2d(x, w, padding=pad)
    dy = torch.randn(pre.shape, generator=g)
    (dw,) = torch.autograd.grad(pre, w, dy)
    got = wgrad_v2(x, dy, pad)
    assert float((got - dw).abs().max() / dw.abs().max()) < 1e-5
