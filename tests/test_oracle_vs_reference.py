"""Live check of the CPU oracle against the UNMODIFIED reference (models.py / utils.py imported from /root/reference with
stub `fcn` / `gdown` modules, oracle/ref_import.py) on inputs that are NOT among the committed golden vectors.  Runs in
the dev container only; on a machine without the reference checkout (the GPU box) it is skipped — the golden-vector test
(test_oracle_golden.py) is the portable pin."""
import numpy as np
import pytest
import torch

from oracle import ref_import, szn_oracle as O

pytestmark = pytest.mark.skipif(ref_import.reference_root() is None, reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.load_reference()


def rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("H,W,D,seed", [(24, 31, 5, 101), (50, 34, 10, 102)])
def test_forward_losses_gradients_and_labels(ref, H, W, D, seed):
    M, U = ref
    C = 21
    params = O.init_params(D, seed)
    model = M.FCN32s(D)
    model.load_state_dict(params, strict=True)
    model.eval()
    x, lab, table = O.synth_batch(1, H, W, C, D, seed=seed, block=8)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    # forward, both heads (models.py:114-160)
    f_ref, s_ref = model(x, mode="both")
    f, s = O.forward(x, pr, "both")
    assert rel(f, f_ref) < 1e-6 and rel(s, s_ref) < 1e-6
    te = O.target_embed_from_labels(lab, table)
    # the dataset-side gather (pascal_dataset.py:122-128) restated
    emb = torch.nn.Embedding(C, D)
    emb.weight.data.copy_(table)
    lbl0 = lab[0].clone()
    lbl0[lbl0 == -1] = 0
    assert torch.equal(te[0], emb(lbl0).data.permute(2, 0, 1))
    # losses (utils.py:19-102) and their gradients through the whole net
    pairs = [(U.cosine_loss(f_ref, lab, te), O.cosine_loss(f, lab, te)),
             (U.mse_loss(f_ref, lab, te), O.mse_loss(f, lab, te)),
             (U.cross_entropy2d(f_ref, lab.clamp(max=D - 1), size_average=False),
              O.cross_entropy2d(f, lab.clamp(max=D - 1), size_average=False)),
             (U.cross_entropy2d(s_ref, O.seenmask_target(lab, [3, 7], C), size_average=True),
              O.cross_entropy2d(s, O.seenmask_target(lab, [3, 7], C), size_average=True))]
    for i, (l_ref, l) in enumerate(pairs):
        assert abs(l.item() - l_ref.item()) <= 1e-5 * max(1.0, abs(l_ref.item())), i
        model.zero_grad()
        for v in pr.values():
            v.grad = None
        l_ref.backward(retain_graph=True)
        l.backward(retain_graph=True)
        named = dict(model.named_parameters())
        for n in ("score_fr.weight", "seenmask_score.weight", "fc7.bias", "conv3_2.weight", "conv1_1.weight"):
            g_ref, g = named[n].grad, pr[n].grad
            if g_ref is None or float(g_ref.abs().max()) == 0.0:
                assert g is None or float(g.abs().max()) == 0.0, (i, n)
                continue
            assert rel(g, g_ref) < 2e-3, (i, n, rel(g, g_ref))
    # inference (utils.py:159-205) on the SAME score tensor: labels are exact
    assert (O.infer_lbl(f_ref.detach(), table) == U.infer_lbl(f_ref.detach(), table)).all()
    seen_tab, unseen_tab = O.split_tables(table, [3, 7, 15])
    assert (O.infer_lbl_forced_unseen(f_ref.detach(), lab, seen_tab, unseen_tab, [3, 7, 15])
            == U.infer_lbl_forced_unseen(f_ref.detach(), lab, seen_tab, unseen_tab, [3, 7, 15])).all()
    assert (O.infer_lbl_szn(f_ref.detach(), s_ref.detach(), seen_tab, unseen_tab)
            == U.infer_lbl_szn(f_ref.detach(), s_ref.detach(), seen_tab, unseen_tab)).all()


def test_seenmask_target_and_tables_follow_the_trainers():
    """trainer_seenmask.py:53-56 and trainer_fcn.py:44,55-64, restated inline with the reference's numpy calls."""
    C, unseen = 21, [3, 7, 15]
    _, lab, table = O.synth_batch(2, 30, 22, C, 5, seed=7, block=4)
    seen = [x for x in range(C) if x not in unseen]
    want = np.in1d(lab.numpy().ravel(), seen).reshape(lab.shape).astype(int)
    assert (O.seenmask_target(lab, unseen, C).numpy() == want).all()
    arr = table.numpy()
    seen_arr, unseen_arr = np.zeros(arr.shape), np.zeros(arr.shape)
    seen_arr[seen, :] = arr[seen, :]
    unseen_arr[unseen, :] = arr[unseen, :]
    st, ut = O.split_tables(table, unseen)
    assert np.array_equal(st.numpy(), seen_arr.astype(np.float32)) and np.array_equal(ut.numpy(), unseen_arr.astype(np.float32))


def test_metrics_follow_utils(ref):
    _, U = ref
    g = np.random.RandomState(3)
    lt = g.randint(-1, 21, size=(3, 20, 20))
    lp = g.randint(0, 21, size=(3, 20, 20))
    want = U.label_accuracy_score(lt, lp, 21)
    hist = sum(O.fast_hist(a.flatten(), b.flatten(), 21) for a, b in zip(lt, lp))
    assert np.allclose(O.hist_to_metrics(hist), want, rtol=1e-12, equal_nan=True)


def test_reference_trainer_forward_szn_unmodified(ref, tmp_path, monkeypatch):
    """The UNMODIFIED ``trainer_fcn.Trainer`` (stub ``pytz``; ``forward_szn`` is the trainer entry point that still runs on
    torch 2.x — ``forward`` indexes a 0-dim loss, ``trainer_fcn.py:107``) with the shipped 20-d PASCAL table: its table
    setup (``trainer_fcn.py:44-64``) and its model -> loss -> stitched-labels sequence (``:123-143``) against the oracle
    and against the product's host-side table split."""
    import datetime
    import importlib.util
    import os
    import sys
    import types
    M, U = ref
    root = ref_import.reference_root()
    pytz = types.ModuleType("pytz")
    pytz.timezone = lambda name: datetime.timezone(datetime.timedelta(hours=-5))
    monkeypatch.setitem(sys.modules, "pytz", pytz)
    monkeypatch.setitem(sys.modules, "utils", U)     # the trainer's `import utils` / `import vis_utils` (generic names)
    monkeypatch.syspath_prepend(root)
    monkeypatch.chdir(root)                          # embedding pickles are opened by relative path (trainer_fcn.py:49)
    before = set(sys.modules)
    try:
        spec = importlib.util.spec_from_file_location("szn_reference_trainer_fcn", os.path.join(root, "trainer_fcn.py"))
        T = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(T)
        D, C, H, W, unseen = 20, 21, 37, 45, [3, 7, 15]
        params = O.init_params(D, seed=55)
        model = M.FCN32s(D)
        model.load_state_dict(params)
        model.eval()
        loader = types.SimpleNamespace(dataset=types.SimpleNamespace(class_names=["c%d" % i for i in range(C)]))
        tr = T.Trainer(cuda=False, model=model, optimizer=None, train_loader=loader, val_loader=loader,
                       log_dir=str(tmp_path), dataset="pascal", max_epoch=1, tb_writer=None, pixel_embeddings=D,
                       loss_func="cos", unseen=unseen, val_unseen=[15], label_names=None, forced_unseen=False)
        table = tr.embeddings.data.clone()
        assert table.shape == (C, D) and tr.n_class == C and tr.seen == [c for c in range(C) if c not in unseen]
        # table setup: oracle restatement and the product's host-side helper
        st, ut = O.split_tables(table, unseen)
        assert torch.equal(tr.seen_embeddings.data, st) and torch.equal(tr.unseen_embeddings.data, ut)
        from zeroshotsemanticsegmentation_b200 import utils as PU
        ps, pu = PU.split_embeddings(table, unseen)
        assert torch.equal(ps, st) and torch.equal(pu, ut)
        # the call sequence
        x, lab, _ = O.synth_batch(1, H, W, C, D, seed=55, block=8)
        te = O.target_embed_from_labels(lab, table)
        fcn_score, loss, lbl_pred, lbl_true = tr.forward_szn(x, (lab, te))
        assert torch.equal(lbl_true, lab)
        p = {k: v.clone() for k, v in params.items()}
        with torch.no_grad():
            f, s = O.forward(x, p, "both")
        assert rel(f, fcn_score) < 1e-6
        assert abs(O.cosine_loss(f, lab, te).item() - loss.item()) < 1e-6
        with torch.no_grad():
            f_ref, s_ref = model(x, mode="both")
        assert (O.infer_lbl_szn(f_ref, s_ref, st, ut) == lbl_pred).all()
        assert (tr.train_log_headers[2], tr.val_log_headers[-1]) == ("train/loss", "elapsed_time") and len(tr.val_log_headers) == 16
    finally:
        for k in set(sys.modules) - before:
            del sys.modules[k]


def test_reference_seenmask_trainer_forward_unmodified(ref, tmp_path, monkeypatch):
    """The UNMODIFIED ``trainer_seenmask.Trainer.forward`` (``trainer_seenmask.py:50-70``): binary target from the label
    map, ``mode='seenmask'``, mean cross entropy, arg-max of the two channels — against the oracle and the product's
    device-agnostic ``utils.seenmask_target``."""
    import datetime
    import importlib.util
    import os
    import sys
    import types
    M, U = ref
    root = ref_import.reference_root()
    pytz = types.ModuleType("pytz")
    pytz.timezone = lambda name: datetime.timezone(datetime.timedelta(hours=-5))
    monkeypatch.setitem(sys.modules, "pytz", pytz)
    monkeypatch.setitem(sys.modules, "utils", U)
    monkeypatch.syspath_prepend(root)
    before = set(sys.modules)
    try:
        spec = importlib.util.spec_from_file_location("szn_reference_trainer_seenmask",
                                                      os.path.join(root, "trainer_seenmask.py"))
        T = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(T)
        D, C, H, W, unseen = 5, 21, 30, 41, [0, 12]
        params = O.init_params(D, seed=56)
        model = M.FCN32s(D)
        model.load_state_dict(params)
        model.eval()
        loader = types.SimpleNamespace(dataset=types.SimpleNamespace(class_names=["c%d" % i for i in range(C)]))
        tr = T.Trainer(cuda=False, model=model, optimizer=None, train_loader=loader, val_loader=loader,
                       log_dir=str(tmp_path), dataset="pascal", max_epoch=1, tb_writer=None, checkpoint={}, unseen=unseen)
        x, lab, table = O.synth_batch(2, H, W, C, D, seed=56, block=8)
        score, loss, lbl_pred, lbl_true = tr.forward(x, (lab, O.target_embed_from_labels(lab, table)))
        tgt = O.seenmask_target(lab, unseen, C)
        assert torch.equal(lbl_true, tgt)
        from zeroshotsemanticsegmentation_b200 import utils as PU
        assert torch.equal(PU.seenmask_target(lab, unseen, C), tgt)   # -1 -> 0 ("unseen"), as upstream
        with torch.no_grad():
            s = O.forward(x, {k: v.clone() for k, v in params.items()}, "seenmask")
        assert rel(s, score) < 1e-6
        assert abs(O.cross_entropy2d(s, tgt, size_average=True).item() - loss.item()) < 1e-6
        assert (score.detach().max(1)[1].numpy() == lbl_pred).all()
    finally:
        for k in set(sys.modules) - before:
            del sys.modules[k]
