"""The callers of the hot path on the CUDA path: the FCN-phase trainer (trainer_fcn.py:83-158,
181-292) and the seen-mask-phase trainer (trainer_seenmask.py:50-101, train.py:166-175), driven through their
reference-shaped API on small synthetic loaders and checked against the CPU oracle + torch.optim."""
import csv
import os
import types

import numpy as np
import pytest
import torch

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
D, C = 20, 21
UNSEEN, VAL_UNSEEN = [3, 7, 15], [15]


class Loader(list):
    """A list of batches that looks like the reference's DataLoader to the trainers (``.dataset.class_names``)."""

    def __init__(self, batches, n_class=C):
        super().__init__(batches)
        self.dataset = types.SimpleNamespace(class_names=["class%d" % i for i in range(n_class)])


def sample(H, W, seed):
    x, lab, _ = O.synth_batch(1, H, W, C, D, seed=seed, block=8)
    return x[0], lab[0]


def table():
    return O.synth_batch(1, 8, 8, C, D, seed=5)[2]


def build_model(seed, train=False):
    import zeroshotsemanticsegmentation_b200 as szn
    params = O.init_params(D, seed=seed)
    m = szn.FCN32s(n_class=D)
    m.load_state_dict(params)
    m = m.to(DEV)
    return (m.train() if train else m.eval()), params


def agreement(a, b):
    return float((np.asarray(a) == np.asarray(b)).mean())


def fcn_trainer(m, train, val, log_dir, **kw):
    from zeroshotsemanticsegmentation_b200 import trainer as T
    from zeroshotsemanticsegmentation_b200.optim import FusedSGD
    ps = [p for n, p in m.named_parameters() if "upscore" not in n and not n.startswith("seenmask")]
    opt = FusedSGD(ps, lr=1e-4, momentum=0.99, weight_decay=5e-4)
    args = dict(pixel_embeddings=D, loss_func="cos", unseen=UNSEEN, val_unseen=VAL_UNSEEN, embed_arr=table().numpy())
    args.update(kw)
    tr = T.Trainer(True, m, opt, train, val, log_dir, "pascal", 1, None, **args)
    tr.verbose = False
    return tr


def test_fcn_trainer_forward_contract_and_parity(tmp_path):
    from zeroshotsemanticsegmentation_b200 import trainer as T
    m, params = build_model(11)
    tab = table()
    data, target = T.collate_padded([sample(40, 56, 1), sample(33, 48, 2)])  # variable-size images, B = 2
    assert data.shape == (2, 3, 40, 56)
    tr = fcn_trainer(m, Loader([(data, target)]), Loader([(data, target)]), str(tmp_path))
    assert tr.n_class == C and tr.seen == [c for c in range(C) if c not in UNSEEN]
    score, loss, lbl_pred, lbl_true = tr.forward(data, target)
    # reference return contract (trainer_fcn.py:111-120)
    assert score.is_cuda and score.shape == (2, D, 40, 56) and score.requires_grad
    assert isinstance(lbl_pred, np.ndarray) and lbl_pred.dtype == np.int64 and lbl_pred.shape == (2, 40, 56)
    assert isinstance(lbl_true, torch.Tensor) and not lbl_true.is_cuda and torch.equal(lbl_true, target)
    with torch.no_grad():
        f, s = O.forward(data, params, "both")
    loss_ref = O.cosine_loss(f, target, O.target_embed_from_labels(target, tab))
    assert abs(loss.item() - loss_ref.item()) < 1e-4
    assert agreement(lbl_pred, O.infer_lbl(f, tab)) > 0.995   # TF32 score vs fp32 oracle score: near-ties may move
    # a reference-style loader item (lbl, lbl_vec) gives the same loss as the on-device table gather
    _, loss2, _, _ = tr.forward(data, (target, O.target_embed_from_labels(target, tab)))
    assert abs(loss2.item() - loss.item()) < 1e-6
    # forced_unseen inference (trainer_fcn.py:112-113) and the SZN stitch (trainer_fcn.py:123-143)
    seen_tab, unseen_tab = O.split_tables(tab, UNSEEN)
    assert torch.equal(tr.seen_embeddings.cpu(), seen_tab) and torch.equal(tr.unseen_embeddings.cpu(), unseen_tab)
    tr.forced_unseen = True
    _, _, lp_forced, _ = tr.forward(data, target)
    assert agreement(lp_forced, O.infer_lbl_forced_unseen(f, target, seen_tab, unseen_tab, UNSEEN)) > 0.995
    tr.forced_unseen = False
    fs, loss3, lp_szn, lt3 = tr.forward_szn(data, target)
    assert abs(loss3.item() - loss_ref.item()) < 1e-4 and isinstance(lp_szn, np.ndarray)
    assert agreement(lp_szn, O.infer_lbl_szn(f, s, seen_tab, unseen_tab)) > 0.99


def nan_equal(a, b):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=1e-12, atol=0, equal_nan=True)


def test_fcn_trainer_epoch_validation_logs_and_checkpoint(tmp_path):
    import zeroshotsemanticsegmentation_b200 as szn
    from zeroshotsemanticsegmentation_b200 import trainer as T
    from zeroshotsemanticsegmentation_b200.optim import FusedSGD
    U = szn.utils
    m, params = build_model(12)
    batches = [T.collate_padded([sample(40, 56, 3), sample(36, 50, 4)]), T.collate_padded([sample(33, 48, 5)])]
    log_dir = str(tmp_path / "run")
    tr = fcn_trainer(m, Loader(batches), Loader(batches), log_dir)
    before = m.score_fr.weight.detach().clone()
    tr.train()  # max_epoch = 1: train_epoch + validate (trainer_fcn.py:294-299)
    assert tr.iteration == 2 and tr.epoch == 0 and not m.training
    assert not torch.equal(before, m.score_fr.weight.detach())
    rows = list(csv.reader(open(os.path.join(log_dir, "train_log.csv"))))
    assert rows[0] == T.TRAIN_HEADERS and len(rows) == 3 and all(len(r) == 8 for r in rows)
    assert all(np.isfinite(float(r[2])) for r in rows[1:])
    vrows = list(csv.reader(open(os.path.join(log_dir, "val_log.csv"))))
    assert len(vrows[0]) == 16 and len(vrows) == 2 and len(vrows[1]) == 16  # all / seen / unseen metric groups
    # validation metrics == the reference's host formulas (utils.py:104-154) on the label maps forward() returns
    val_loss, (all_m, seen_m, unseen_m) = tr.validate()
    lts, lps, losses = [], [], []
    for data, target in batches:
        with torch.no_grad():
            _, loss, lp, lt = tr.forward(data, target)
        losses.append(loss.item())
        lts += [t for t in lt.numpy()]
        lps += [p for p in lp]
    want = U.label_accuracy_score(lts, lps, C, unseen=VAL_UNSEEN)
    assert nan_equal(all_m, want[0]) and nan_equal(seen_m, want[1]) and nan_equal(unseen_m, want[2])
    assert abs(val_loss - np.mean(losses)) < 1e-6
    # checkpoint in the reference's dictionary format; resume restores model and optimizer (train.py:110-116,135-136)
    path = os.path.join(log_dir, "checkpoint")
    assert os.path.isfile(path)
    if tr.best_mean_iu > 0:
        assert os.path.isfile(os.path.join(log_dir, "best"))
    m2 = szn.FCN32s(n_class=D).to(DEV)
    ps2 = [p for n, p in m2.named_parameters() if "upscore" not in n and not n.startswith("seenmask")]
    opt2 = FusedSGD(ps2, lr=1e-4, momentum=0.99, weight_decay=5e-4)
    ck = T.load_checkpoint(path, m2, opt2)
    assert ck["arch"] == "FCN32s" and ck["iteration"] == 2
    for (n1, a), (n2, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert n1 == n2 and torch.equal(a, b), n1
    s1, s2 = tr.optim.state_dict()["state"], opt2.state_dict()["state"]
    assert len(s1) == len(s2) > 0
    assert all(torch.equal(s1[k]["momentum_buffer"], s2[k]["momentum_buffer"]) for k in s1)
    # the stitched validation of phase 3 (train.py:194-196) runs on the same loader
    tr.validate(both_fcn_and_seenmask=True)


def test_seenmask_phase_matches_oracle_and_torch_adam(tmp_path):
    from zeroshotsemanticsegmentation_b200 import trainer as T
    from zeroshotsemanticsegmentation_b200.optim import FusedAdam
    m, params = build_model(13, train=True)
    g = torch.Generator().manual_seed(3)
    masks = ((torch.rand(1, 4096, generator=g) < 0.5).float(), (torch.rand(1, 4096, generator=g) < 0.5).float())
    m._forced_drop_masks = masks  # Dropout2d is live in train_epoch: pin its masks so that the oracle can replay it
    batches = [T.collate_padded([sample(40, 56, s)]) for s in (6, 7, 8)]
    lr = 1e-3
    head = T.freeze_for_seenmask(m)                       # train.py:166-171
    assert [tuple(p.shape) for p in head] == [(2, 4096, 1, 1), (2,), (2, 2, 64, 64)]
    opt = FusedAdam([{"params": head}], lr=lr)            # train.py:174-175
    tr = T.SeenmaskTrainer(True, m, opt, Loader(batches), Loader(batches[:1]), str(tmp_path), "pascal", 1, None,
                           checkpoint={"epoch": 7}, unseen=UNSEEN)
    tr.verbose = False
    # reference return contract (trainer_seenmask.py:50-70)
    score, loss, lbl_pred, lbl_true = tr.forward(*batches[0])
    assert score.shape == (1, 2, 40, 56) and isinstance(lbl_pred, np.ndarray) and lbl_pred.shape == (1, 40, 56)
    assert torch.equal(lbl_true, O.seenmask_target(batches[0][1], UNSEEN, C))
    tr.train_epoch()
    assert tr.iteration == 3
    # frozen trunk: untouched, no gradients (train.py:166-167)
    assert torch.equal(m.conv1_1.weight.detach().cpu(), params["conv1_1.weight"]) and m.conv1_1.weight.grad is None
    assert torch.equal(m.score_fr.weight.detach().cpu(), params["score_fr.weight"]) and m.fc6.weight.grad is None
    rows = list(csv.reader(open(os.path.join(str(tmp_path), "seenmask_train_log.csv"))))
    losses = [float(r[2]) for r in rows[1:]]
    # ---- oracle replay with torch.optim.Adam ----
    names = ["seenmask_score.weight", "seenmask_score.bias", "seenmask_upscore.weight"]
    pr = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
    ref_opt = torch.optim.Adam([pr[n] for n in names], lr=lr)
    ref_losses = []
    for data, target in batches:
        s = O.forward(data, pr, "seenmask", drop_masks=masks)
        l = O.cross_entropy2d(s, O.seenmask_target(target, UNSEEN, C), size_average=True)
        ref_opt.zero_grad()
        l.backward()
        ref_opt.step()
        ref_losses.append(l.item())
    print("losses", losses, ref_losses)
    assert np.allclose(losses, ref_losses, rtol=5e-3, atol=1e-4)
    sd = m.state_dict()
    for n in names:
        a, b = sd[n].cpu(), pr[n].detach()
        upd, err = (b - params[n]).norm().item(), (a - b).norm().item()
        print(n, "update norm %.3e  error %.3e" % (upd, err))
        assert upd > 0 and err < 0.2 * upd
    # validate() rewrites the checkpoint handed in by train.py:177-181 as `best` with the new weights
    val_loss, metrics = tr.validate()
    assert np.isfinite(val_loss) and len(metrics) == 4
    best = torch.load(os.path.join(str(tmp_path), "best"), weights_only=False)
    assert best["epoch"] == 7 and torch.equal(best["model_state_dict"]["seenmask_score.weight"].cpu(),
                                              sd["seenmask_score.weight"].cpu())


class _FakeReducer:
    """Stands in for ddp.GradientAllReduce on one GPU: behaves as if a second rank had contributed the same
    {sum, n_valid} to the loss accumulator."""

    def __init__(self):
        self.seen = []

    def accum_hook(self, accum):
        self.seen.append(accum.clone())
        accum.mul_(2.0)


def test_cross_entropy_accum_hook_uses_the_global_count():
    """utils.py:46-47 under data parallelism: the mean divides by the valid pixels of ALL ranks."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = torch.Generator().manual_seed(4)
    score = torch.randn(2, 5, 9, 11, generator=g).to(DEV)
    tgt = torch.randint(-1, 5, (2, 9, 11), generator=g).to(DEV)
    n_valid = int((tgt >= 0).sum())
    for size_average in (True, False):
        s0 = score.clone().requires_grad_(True)
        base = U.cross_entropy2d(s0, tgt, size_average=size_average)
        base.backward()
        red = _FakeReducer()
        s1 = score.clone().requires_grad_(True)
        loss = U.cross_entropy2d(s1, tgt, size_average=size_average, accum_hook=red.accum_hook)
        loss.backward()
        acc = red.seen[0].cpu()
        assert int(acc[1]) == n_valid
        if size_average:
            assert abs(float(acc[0]) / n_valid - base.item()) < 1e-5
            assert abs(loss.item() - base.item()) < 1e-6          # 2*sum / 2*N
            assert torch.allclose(s1.grad, s0.grad * 0.5, rtol=1e-6, atol=1e-12)   # each rank's share of the global mean
        else:
            assert abs(loss.item() - 2 * base.item()) < 1e-4 * abs(base.item())  # the global sum
            assert torch.allclose(s1.grad, s0.grad, rtol=1e-6, atol=1e-12)       # a sum's gradient does not depend on N


def test_trainers_hand_the_reducer_hook_to_their_losses(tmp_path):
    from zeroshotsemanticsegmentation_b200 import trainer as T
    from zeroshotsemanticsegmentation_b200.optim import FusedAdam
    m, _ = build_model(14)
    batch = T.collate_padded([sample(40, 56, 9)])
    red = _FakeReducer()
    tr = fcn_trainer(m, Loader([batch]), Loader([batch]), None, reducer=red)
    _, loss, _, _ = tr.forward(*batch)
    tr0 = fcn_trainer(m, Loader([batch]), Loader([batch]), None)
    _, loss0, _, _ = tr0.forward(*batch)
    assert len(red.seen) == 1 and int(red.seen[0][1]) == int((batch[1] >= 0).sum())
    assert abs(loss.item() - loss0.item()) < 1e-6  # (2N - 2*sum) / 2N
    head = T.freeze_for_seenmask(m)
    sm = T.SeenmaskTrainer(True, m, FusedAdam([{"params": head}], lr=1e-3), Loader([batch]), Loader([batch]), None, "pascal",
                           1, None, unseen=UNSEEN, reducer=red)
    _, l1, _, _ = sm.forward(*batch)
    l1.backward()
    g1 = m.seenmask_score.weight.grad.clone()
    m.zero_grad()
    sm.reducer, sm._accum_hook = None, None
    _, l0, _, _ = sm.forward(*batch)
    l0.backward()
    assert len(red.seen) == 2 and abs(l1.item() - l0.item()) < 1e-6
    assert torch.allclose(g1, m.seenmask_score.weight.grad * 0.5, rtol=1e-4, atol=1e-10)
