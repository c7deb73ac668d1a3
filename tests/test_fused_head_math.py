"""The algebra of the round-2 head (tools/fused_head_math.py: loss, labels and d s17 from 17x17-sized quantities, no
(B,D,H,W) tensor) against the oracle's materialised path (models.py:94,146-147 + utils.py:75-102,159-185).  CPU only."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from oracle import szn_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fused_head_math as FH  # noqa: E402


@pytest.mark.parametrize("kind", [0, 1], ids=["cos", "mse"])
@pytest.mark.parametrize("B,D,C,H,W,hs,ws", [(2, 6, 7, 40, 56, 2, 3), (1, 5, 9, 70, 33, 3, 2), (1, 4, 3, 5, 3, 1, 1)])
def test_fused_head_equals_materialised_path(B, D, C, H, W, hs, ws, kind):
    g = torch.Generator().manual_seed(H)
    s17 = torch.randn(B, D, hs, ws, generator=g, dtype=torch.float64, requires_grad=True)
    _, lab, table = O.synth_batch(B, H, W, C, D, seed=H, block=4, ignore_frac=0.1)
    table = table.double()
    # materialised: dense diagonal transposed conv + crop, then the oracle's loss / labels
    w = O.upsampling_weight(D, D).double()
    up = F.conv_transpose2d(s17, w, stride=32)[:, :, 19:19 + H, 19:19 + W]
    assert up.shape == (B, D, H, W)
    # the tap matrix IS the transposed conv + crop
    P = FH.tap_matrix(H, W, hs, ws)
    up2 = torch.einsum("pk,bdk->bdp", P, s17.detach().reshape(B, D, hs * ws)).reshape(B, D, H, W)
    assert torch.allclose(up2, up.detach(), atol=1e-12)
    loss_fn = O.cosine_loss if kind == 0 else O.mse_loss
    loss_ref = loss_fn(up, lab, O.target_embed_from_labels(lab, table))
    (g_ref,) = torch.autograd.grad(loss_ref, s17)
    lbl_ref = O.infer_lbl(up.detach(), table)
    loss, labels, ds17 = FH.fused_cosine_head(s17.detach(), lab, table, kind)
    assert abs(loss.item() - loss_ref.item()) < 1e-10 * max(1.0, abs(loss_ref.item()))
    assert torch.allclose(ds17, g_ref, rtol=1e-8, atol=1e-12)
    assert (labels.numpy() == lbl_ref).all()


@pytest.mark.parametrize("kind", [0, 1], ids=["cos", "mse"])
@pytest.mark.parametrize("B,D,C,H,W,hs,ws", [(2, 6, 7, 40, 56, 2, 3), (1, 5, 9, 70, 33, 3, 2), (1, 4, 3, 5, 3, 1, 1)])
def test_kernel_emulation_equals_the_algebra(B, D, C, H, W, hs, ws, kind):
    """tools/fused_head_emulate.py follows csrc/szn_fused_head.cu index for index (cells, taps, Gram pairs, M2 expansion,
    gather around a node); it must reproduce the dense-matrix algebra above."""
    import fused_head_emulate as EM
    g = torch.Generator().manual_seed(H + 1)
    s17 = torch.randn(B, D, hs, ws, generator=g, dtype=torch.float64)
    _, lab, table = O.synth_batch(B, H, W, C, D, seed=H + 1, block=4, ignore_frac=0.1)
    table = table.double()
    table[C // 2] = 0  # a zero row: similarity exactly 0, never a target here
    lab[lab == C // 2] = -1
    loss_a, labels_a, ds_a = FH.fused_cosine_head(s17, lab, table, kind)
    loss_e, labels_e, ds_e = EM.fused_cosine_head(s17.numpy(), lab.numpy(), table.numpy(), kind)
    assert abs(loss_e - loss_a.item()) < 1e-12 * max(1.0, abs(loss_a.item()))
    assert (labels_e == labels_a.numpy()).all()
    assert abs(ds_e - ds_a.numpy()).max() < 1e-12 * max(1.0, abs(ds_a.numpy()).max())
