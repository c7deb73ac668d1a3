"""Edge cases of the hot path on the CUDA path, each against the CPU oracle (= what the reference does there):
all-ignored label maps, the smallest and most ragged image sizes (pad=100 makes every H, W >= 1 legal),
degenerate class tables, out-of-range labels in the metrics, non-contiguous score tensors."""
import math

import numpy as np
import pytest
import torch

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_all_pixels_ignored_is_nan_like_the_reference():
    """N_valid == 0: the reference divides by zero (utils.py:72,101,46-47) and the trainer raises on the NaN
    (trainer_fcn.py:107-108); cross_entropy2d with size_average=False returns 0."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(0)
    score = torch.randn(2, 6, 9, 11, generator=g)
    tab = torch.randn(4, 6, generator=g)
    lab = torch.full((2, 9, 11), -1, dtype=torch.int64)
    te = O.target_embed_from_labels(lab, tab)
    assert math.isnan(O.cosine_loss(score, lab, te).item()) and math.isnan(O.mse_loss(score, lab, te).item())
    sd, ld, td = score.to(DEV), lab.to(DEV), tab.to(DEV)
    assert math.isnan(U.cosine_loss(sd, ld, table=td).item())
    assert math.isnan(U.mse_loss(sd, ld, te.to(DEV)).item())
    assert math.isnan(U.cross_entropy2d(sd, ld, size_average=True).item())
    assert U.cross_entropy2d(sd, ld, size_average=False).item() == 0.0
    assert O.cross_entropy2d(score, lab, size_average=False).item() == 0.0


# (500, 375): a native-size PASCAL portrait image, W % 4 != 0 and W > 256 (the upsample backward refused it)
@pytest.mark.parametrize("H,W,B", [(1, 1, 1), (5, 3, 2), (24, 31, 1), (33, 1, 3), (500, 375, 1)])
def test_smallest_and_ragged_images(H, W, B):
    """Whole path at sizes where the 17x17-style score map degenerates to 1x1 / 1x2 / 2x1, H*W is odd, B is odd."""
    import zeroshotsemanticsegmentation_b200 as szn
    U = szn.utils
    D, C = 5, 7
    params = O.init_params(D, seed=21)
    x, lab, tab = O.synth_batch(B, H, W, C, D, seed=21 + H, block=2, ignore_frac=0.0 if H * W < 4 else 0.2)
    pr = {k: v.clone().requires_grad_("upscore" not in k) for k, v in params.items()}
    f_ref, s_ref = O.forward(x, pr, "both")
    loss_ref = O.mse_loss(f_ref, lab, O.target_embed_from_labels(lab, tab))
    loss_ref.backward()
    m = szn.FCN32s(D)
    m.load_state_dict(params)
    m = m.to(DEV).eval()
    f, s = m(x.to(DEV), mode="both")
    assert f.shape == (B, D, H, W) and s.shape == (B, 2, H, W) and f.is_contiguous()
    assert rel(f.detach().cpu().numpy(), f_ref.detach().numpy()) < 1e-3
    assert rel(s.detach().cpu().numpy(), s_ref.detach().numpy()) < 1e-3
    loss = U.mse_loss(f, lab.to(DEV), table=tab.to(DEV))
    assert abs(loss.item() - loss_ref.item()) < 1e-3 * max(1.0, abs(loss_ref.item()))
    loss.backward()
    assert rel(m.score_fr.weight.grad.cpu().numpy(), pr["score_fr.weight"].grad.numpy()) < 1e-2
    assert rel(m.score_fr.bias.grad.cpu().numpy(), pr["score_fr.bias"].grad.numpy()) < 1e-2
    # labels on the oracle's score tensor: exact
    fd = f_ref.detach()
    assert (U.infer_lbl(fd.to(DEV), tab.to(DEV)) == O.infer_lbl(fd, tab)).all()
    st, ut = O.split_tables(tab, [1, 4])
    sd = s_ref.detach()
    assert (U.infer_lbl_szn(fd.to(DEV), sd.to(DEV), st.to(DEV), ut.to(DEV)) == O.infer_lbl_szn(fd, sd, st, ut)).all()


def test_degenerate_tables_and_ties():
    """All-zero table: every similarity is exactly 0, the first index wins (utils.py:172-180, torch.max tie rule).
    Duplicate rows: the lower index wins.  One class only: label 0 everywhere."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(4)
    for shape in ((1, 8, 16, 16), (2, 33, 5, 7)):  # tensor-core path (H*W % 32 == 0) and the CUDA-core path
        score = torch.randn(shape, generator=g)
        D = shape[1]
        zero = torch.zeros(6, D)
        assert (U.infer_lbl(score.to(DEV), zero.to(DEV)) == 0).all() and (O.infer_lbl(score, zero) == 0).all()
        dup = torch.randn(3, D, generator=g)
        dup = torch.cat([dup, dup], 0)  # rows 3..5 repeat rows 0..2 exactly: exact ties, the lower index wins
        got = U.infer_lbl(score.to(DEV), dup.to(DEV))
        assert got.max() <= 2
        one = torch.randn(1, D, generator=g)
        assert (U.infer_lbl(score.to(DEV), one.to(DEV)) == 0).all()


def test_metrics_ignore_out_of_range_labels():
    """_fast_hist keeps 0 <= label_true < n_class only (utils.py:105)."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    lt = torch.tensor([[[-1, 0, 1, 2, 5, 300, -7, 1]]])
    lp = torch.tensor([[[0, 0, 1, 1, 2, 0, 1, 1]]])
    hist = U.confusion_hist_device(lt.to(DEV), lp.to(DEV), 3).cpu().numpy()[0]
    assert (hist == O.fast_hist(lt.numpy().ravel(), lp.numpy().ravel(), 3)).all() and hist.sum() == 4


def test_non_contiguous_scores_and_explicit_target_embed():
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(8)
    big = torch.randn(2, 12, 10, 14, generator=g)
    tab = torch.randn(9, 6, generator=g)
    lab = torch.randint(-1, 9, (2, 10, 14), generator=g)
    score = big[:, ::2]                                           # strided channel slice
    cl = score.contiguous(memory_format=torch.channels_last)      # NHWC memory behind an NCHW shape
    te = O.target_embed_from_labels(lab, tab)
    want = O.cosine_loss(score, lab, te).item()
    for sc in (score, cl):
        sd = sc.to(DEV)
        if sc is score:
            sd = big.to(DEV)[:, ::2]
        sd.requires_grad_(False)
        assert abs(U.cosine_loss(sd, lab.to(DEV), te.to(DEV)).item() - want) < 1e-5
        assert abs(U.cosine_loss(sd, lab.to(DEV), table=tab.to(DEV)).item() - want) < 1e-5
        assert (U.infer_lbl(sd, tab.to(DEV)) == O.infer_lbl(score.contiguous(), tab)).mean() > 0.99


def test_out_of_range_labels_poison_the_loss_instead_of_reading_out_of_bounds():
    """a label >= the table's rows (an un-remapped 255, a 59- vs 33-class table) or >= the CE channel count used
    to index past the table / score and return a plausible finite number.  torch's embedding / nll_loss device-assert on
    such input; here the loss is NaN (the trainers' NaN guard, trainer_fcn.py:107-108, raises) and backward stays in
    bounds."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(4)
    score = torch.randn(2, 6, 8, 12, generator=g).to(DEV).requires_grad_(True)
    tab = torch.randn(5, 6, generator=g).to(DEV)
    lab = torch.randint(0, 5, (2, 8, 12), generator=g)
    ok = U.cosine_loss(score, lab.to(DEV), table=tab)
    assert math.isfinite(ok.item())
    bad = lab.clone()
    bad[1, 3, 4] = 5      # == rows
    bad[0, 0, 0] = 255    # the raw VOC ignore label, not remapped to -1
    for fn in (U.cosine_loss, U.mse_loss):
        loss = fn(score, bad.to(DEV), table=tab)
        assert math.isnan(loss.item())
        loss.backward()   # must not fault
        torch.cuda.synchronize()
    ce_score = torch.randn(2, 5, 8, 12, generator=g).to(DEV).requires_grad_(True)
    assert math.isfinite(U.cross_entropy2d(ce_score, lab.to(DEV)).item())
    loss = U.cross_entropy2d(ce_score, bad.to(DEV), size_average=True)
    assert math.isnan(loss.item())
    loss.backward()
    torch.cuda.synchronize()
    assert (ce_score.grad[0, :, 0, 0] == 0).all() and (ce_score.grad[1, :, 3, 4] == 0).all()


def test_data_edits_need_invalidate_and_padded_seenmask_batches_ignore_the_padding():
    """(1) parameters edited through .data do not advance the version counter the packed-weight cache keys on:
    invalidate_weight_cache() (also called by load_state_dict / .to() / copy_params_from_vgg16) makes them visible.
    (2) a ragged B=2 seen-mask batch: padded pixels neither enter the mean CE nor get a gradient."""
    import zeroshotsemanticsegmentation_b200 as szn
    from zeroshotsemanticsegmentation_b200 import trainer as T
    U = szn.utils
    D, C = 5, 7
    params = O.init_params(D, seed=3)
    m = szn.FCN32s(D)
    m.load_state_dict(params)
    m = m.to(DEV).eval()
    x, lab, tab = O.synth_batch(1, 24, 31, C, D, seed=9, block=4)
    with torch.no_grad():
        f0 = m(x.to(DEV)).clone()
        m.conv3_2.weight.data.mul_(1.5)      # version counter unchanged: the cached pack is still used
        m.invalidate_weight_cache()
        f1 = m(x.to(DEV)).clone()
    p2 = {k: v.clone() for k, v in params.items()}
    p2["conv3_2.weight"] = p2["conv3_2.weight"] * 1.5
    assert rel(f1.cpu().numpy(), O.forward(x, p2, "fcn").numpy()) < 1e-3 and not torch.equal(f0, f1)
    # ragged seen-mask batch
    a = O.synth_batch(1, 24, 31, C, D, seed=10, block=4)
    b = O.synth_batch(1, 17, 20, C, D, seed=11, block=4)
    data, target = T.collate_padded([(a[0][0], a[1][0]), (b[0][0], b[1][0])])
    unseen = [1, 4]
    t = U.seenmask_target(target.to(DEV), unseen, C)
    s = m(data.to(DEV), mode="seenmask")
    s.retain_grad()
    loss = U.cross_entropy2d(s, t, size_average=True)
    loss.backward()
    # oracle: the same scores, per item on its own unpadded window, normalised by the total valid count
    so = s.detach().cpu()
    tot, cnt = 0.0, 0
    for i, (h, w, item) in enumerate(((24, 31, a), (17, 20, b))):
        ti = O.seenmask_target(item[1], unseen, C)
        tot += O.cross_entropy2d(so[i:i + 1, :, :h, :w], ti, size_average=False).item()
        cnt += int((ti >= 0).sum())
    assert abs(loss.item() - tot / cnt) < 1e-5 * max(1.0, abs(tot / cnt))
    assert (s.grad[1, :, 17:, :] == 0).all() and (s.grad[1, :, :, 20:] == 0).all() and s.grad[1, :, :17, :20].abs().sum() > 0
