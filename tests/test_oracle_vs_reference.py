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
