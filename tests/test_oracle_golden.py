"""Pins the CPU oracle (oracle/szn_oracle.py) to golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import szn_oracle as O


def checksum(p):
    return float(sum(v.double().abs().sum() for k, v in sorted(p.items()) if "upscore" not in k))


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_ce21_small(golden):
    g = golden("ce21_37x53")
    p = {k: v.requires_grad_(True) for k, v in O.init_params(21, int(g["seed"])).items()}
    assert abs(checksum(p) - float(g["param_checksum"])) < 1e-6 * float(g["param_checksum"])
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["target"])
    score = O.forward(x, p, "fcn")
    assert rel(score.detach().numpy(), g["score"]) < 1e-6
    loss = O.cross_entropy2d(score, t)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    loss.backward()
    assert rel(p["score_fr.weight"].grad.numpy(), g["score_fr__weight__grad"]) < 1e-4
    assert rel(p["conv1_1.weight"].grad.numpy(), g["conv1_1__weight__grad"]) < 1e-3
    assert rel(p["fc7.bias"].grad.numpy(), g["fc7__bias__grad"]) < 1e-4
    assert (score.detach().max(1)[1].numpy() == g["lbl"]).all()


def test_cos_batched_equals_reference_loop(golden):
    g = golden("cos_voc20_2x64x96")
    tab = torch.from_numpy(g["table"]).float()
    p = {k: v.requires_grad_(True) for k, v in O.init_params(tab.shape[1], int(g["seed"])).items()}
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["target"]).long()
    f, s = O.forward(x, p, "both")
    assert rel(f.detach().numpy(), g["score"]) < 1e-5
    assert rel(s.detach().numpy(), g["seenmask_score"]) < 1e-5
    te = O.target_embed_from_labels(t, tab)
    loss = O.cosine_loss(f, t, te)
    want = (g["loss_per_sample"] * g["nvalid"]).sum() / g["nvalid"].sum()
    assert abs(loss.item() - want) < 1e-5
    loss.backward()
    assert rel(p["score_fr.weight"].grad.numpy(), g["score_fr__weight__grad"]) < 1e-3
    assert rel(p["conv1_1.weight"].grad.numpy(), g["conv1_1__weight__grad"]) < 2e-3
    # inference on the golden score tensor itself -> exact labels
    fs = torch.from_numpy(g["score"])
    assert (O.infer_lbl(fs, tab) == g["lbl"]).all()
    se, ue = O.split_tables(tab, list(g["unseen"]))
    ss = torch.from_numpy(g["seenmask_score"])
    assert (O.infer_lbl_szn(fs, ss, se, ue) == g["lbl_szn"]).all()
    assert (O.infer_lbl_forced_unseen(fs, t, se, ue, list(g["unseen"])) == g["lbl_forced"]).all()
    for i in range(2):
        smt = O.seenmask_target(t[i:i + 1], list(g["train_unseen"]), tab.shape[0])
        l = O.cross_entropy2d(ss[i:i + 1], smt, size_average=True)
        assert abs(l.item() - g["seenmask_loss"][i]) < 1e-5


def test_mse(golden):
    g = golden("mse_voc20_1x45x70")
    tab = torch.from_numpy(g["table"]).float()
    p = O.init_params(tab.shape[1], int(g["seed"]))
    x, t = torch.from_numpy(g["x"]), torch.from_numpy(g["target"]).long()
    f = O.forward(x, p, "fcn")
    assert rel(f.numpy(), g["score"]) < 1e-5
    loss = O.mse_loss(f, t, O.target_embed_from_labels(t, tab))
    assert abs(loss.item() - g["loss_per_sample"][0]) < 1e-5 * abs(g["loss_per_sample"][0])


@pytest.mark.parametrize("name", ["head_ctx300_24x40", "head_voc20_33x17"])
def test_head_functions(golden, name):
    g = golden(name)
    tab = torch.from_numpy(g["table"]).float()
    t = torch.from_numpy(g["target"]).long()
    te = O.target_embed_from_labels(t, tab)
    s = torch.from_numpy(g["score"]).requires_grad_(True)
    l = O.cosine_loss(s, t, te); l.backward()
    assert abs(l.item() - float(g["cos_loss"])) < 1e-5
    assert rel(s.grad.numpy(), g["cos_grad"]) < 1e-4
    s = torch.from_numpy(g["score"]).requires_grad_(True)
    l = O.mse_loss(s, t, te); l.backward()
    assert abs(l.item() - float(g["mse_loss"])) < 1e-5 * float(g["mse_loss"])
    assert rel(s.grad.numpy(), g["mse_grad"]) < 1e-5
    c = torch.from_numpy(g["ce_score"]).requires_grad_(True)
    l = O.cross_entropy2d(c, torch.from_numpy(g["ce_target"]).long()); l.backward()
    assert abs(l.item() - float(g["ce_loss"])) < 1e-5 * float(g["ce_loss"])
    assert rel(c.grad.numpy(), g["ce_grad"]) < 1e-5
    sm = torch.from_numpy(g["seenmask_score"]).requires_grad_(True)
    smt = O.seenmask_target(t, list(g["unseen"]), tab.shape[0])
    l = O.cross_entropy2d(sm, smt, size_average=True); l.backward()
    assert abs(l.item() - float(g["sm_loss"])) < 1e-5
    assert rel(sm.grad.numpy(), g["sm_grad"]) < 1e-5
    sc = torch.from_numpy(g["score"])
    se, ue = O.split_tables(tab, list(g["unseen"]))
    assert (O.infer_lbl(sc, tab) == g["lbl"]).all()
    assert (O.infer_lbl(sc, se) == g["lbl_seen_only"]).all()
    assert (O.infer_lbl_szn(sc, sm.detach(), se, ue) == g["lbl_szn"]).all()
    assert (O.infer_lbl_forced_unseen(sc, t, se, ue, list(g["unseen"])) == g["lbl_forced"]).all()
    # the zero-row rule: a zeroed class wins where every live cosine is negative (utils.py:175)
    assert np.isin(g["lbl_seen_only"][0, 0], list(g["unseen"])).any()


def test_bilinear_filter_matches_reference_values():
    f = O.bilinear_filter(64)
    assert abs(f.sum().item() - 1024.0) < 1e-3
    assert abs(f[0, 0].item() - (1 / 64) ** 2) < 1e-9
    assert abs(f[31, 31].item() - (1 - 0.5 / 32) ** 2) < 1e-7
