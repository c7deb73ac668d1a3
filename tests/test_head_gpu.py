"""Losses and inference (GPU) against the golden vectors produced by the unmodified reference and
against the CPU oracle on seeded inputs; through the public drop-in API (zeroshotsemanticsegmentation_b200.utils)."""
import numpy as np
import pytest
import torch

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("name", ["head_ctx300_24x40", "head_voc20_33x17"])
def test_losses_and_labels_vs_reference_golden(golden, name):
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden(name)
    tab = torch.from_numpy(g["table"]).float()
    t = torch.from_numpy(g["target"]).long()
    te = O.target_embed_from_labels(t, tab)
    for explicit in (True, False):  # explicit target_embed tensor (reference call) and table gather
        kw = dict(target_embed=te.to(DEV)) if explicit else dict(table=tab.to(DEV))
        s = torch.from_numpy(g["score"]).to(DEV).requires_grad_(True)
        l = U.cosine_loss(s, t.to(DEV), **kw)
        l.backward()
        assert abs(l.item() - float(g["cos_loss"])) < 1e-5
        assert rel(s.grad.cpu().numpy(), g["cos_grad"]) < 1e-4
        s = torch.from_numpy(g["score"]).to(DEV).requires_grad_(True)
        l = U.mse_loss(s, t.to(DEV), **kw)
        l.backward()
        assert abs(l.item() - float(g["mse_loss"])) < 1e-5 * float(g["mse_loss"])
        assert rel(s.grad.cpu().numpy(), g["mse_grad"]) < 1e-5
    c = torch.from_numpy(g["ce_score"]).to(DEV).requires_grad_(True)
    l = U.cross_entropy2d(c, torch.from_numpy(g["ce_target"]).long().to(DEV))
    l.backward()
    assert abs(l.item() - float(g["ce_loss"])) < 1e-5 * float(g["ce_loss"])
    assert rel(c.grad.cpu().numpy(), g["ce_grad"]) < 1e-5
    sm = torch.from_numpy(g["seenmask_score"]).to(DEV).requires_grad_(True)
    smt = O.seenmask_target(t, list(g["unseen"]), tab.shape[0])
    l = U.cross_entropy2d(sm, smt.to(DEV), size_average=True)
    l.backward()
    assert abs(l.item() - float(g["sm_loss"])) < 1e-5
    assert rel(sm.grad.cpu().numpy(), g["sm_grad"]) < 1e-5
    # inference: labels must be bit-exact on identical inputs
    sc = torch.from_numpy(g["score"]).to(DEV)
    se, ue = O.split_tables(tab, list(g["unseen"]))
    lbl = U.infer_lbl(sc, tab.to(DEV))
    assert lbl.dtype == np.int64 and lbl.shape == g["lbl"].shape
    assert (lbl == g["lbl"]).all()
    assert (U.infer_lbl(sc, se.to(DEV)) == g["lbl_seen_only"]).all()
    assert (U.infer_lbl_szn(sc, sm.detach(), se.to(DEV), ue.to(DEV)) == g["lbl_szn"]).all()
    assert (U.infer_lbl_forced_unseen(sc, t.to(DEV), se.to(DEV), ue.to(DEV), list(g["unseen"])) == g["lbl_forced"]).all()


def test_batched_loss_equals_per_sample_reference_loop(golden):
    """n > 1: (sum_i N_i * loss_i) / sum_i N_i of the reference run per sample."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("cos_voc20_2x64x96")
    tab = torch.from_numpy(g["table"]).float()
    t = torch.from_numpy(g["target"]).long()
    s = torch.from_numpy(g["score"]).to(DEV)
    l = U.cosine_loss(s, t.to(DEV), table=tab.to(DEV))
    want = (g["loss_per_sample"] * g["nvalid"]).sum() / g["nvalid"].sum()
    assert abs(l.item() - want) < 1e-5
    assert (U.infer_lbl(s, tab.to(DEV)) == g["lbl"]).all()
    se, ue = O.split_tables(tab, list(g["unseen"]))
    ss = torch.from_numpy(g["seenmask_score"]).to(DEV)
    assert (U.infer_lbl_szn(s, ss, se.to(DEV), ue.to(DEV)) == g["lbl_szn"]).all()


@pytest.mark.parametrize("shape", [(2, 300, 64, 64, 59), (1, 1024, 32, 48, 256), (3, 20, 17, 9, 33)])
def test_argmax_vs_oracle_seeded(shape):
    """Bit-exact labels wherever the oracle's own top-2 margin exceeds fp32 summation noise."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    n, D, h, w, C = shape
    g = torch.Generator().manual_seed(11)
    score = torch.randn(n, D, h, w, generator=g)
    tab = torch.randn(C, D, generator=g)
    tab[C // 2] = 0  # a zero row (an "excluded" class): similarity exactly 0
    ref = O.infer_lbl(score, tab)
    got = U.infer_lbl(score.to(DEV), tab.to(DEV))
    mism = got != ref
    if mism.any():
        # allowed only at numerical near-ties of the oracle itself
        sn = score / score.norm(dim=1, keepdim=True)
        en = tab.norm(dim=1).clone(); en[en == 0] = 1
        sim = torch.einsum("ndhw,cd->nchw", sn, tab / en[:, None])
        top2 = sim.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]).numpy()
        assert (margin[mism] < 1e-5).all(), "label mismatch away from a near-tie"
    assert mism.mean() < 1e-4


@pytest.mark.parametrize("n_class,unseen", [(21, [3, 17]), (59, list(range(49, 59))), (33, None), (256, [0, 255])])
def test_device_metrics_equal_reference_formulas(n_class, unseen):
    """label_accuracy_score on device-resident labels: identical histograms, hence identical scores (utils.py:104-154)."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(17)
    lt = torch.randint(-1, n_class, (3, 61, 47), generator=g)
    lp = torch.randint(0, n_class, (3, 61, 47), generator=g)
    lp[lt >= 0] = torch.where(torch.rand(lt.shape, generator=g) < 0.6, lt, lp)[lt >= 0]  # mostly right, like a trained net
    hist = U.confusion_hist_device(lt.to(DEV), lp.to(DEV), n_class, unseen).cpu().numpy()
    want_all = O.fast_hist(lt.numpy().ravel(), lp.numpy().ravel(), n_class)
    assert (hist[0] == want_all).all()
    got = U.label_accuracy_score(lt.to(DEV), lp.to(DEV), n_class, unseen)
    ref = U.label_accuracy_score([a for a in lt.numpy()], [a for a in lp.numpy()], n_class, unseen)  # numpy path = reference code
    np.testing.assert_allclose(np.array(got, dtype=np.float64), np.array(ref, dtype=np.float64), rtol=0, atol=0)
    if unseen:
        assert (hist[1] + hist[2] == hist[0]).all()
    assert np.allclose(got[0] if unseen else got, O.hist_to_metrics(want_all), equal_nan=True)
