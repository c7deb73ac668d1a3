"""Host logic of the trainer layer that needs no GPU: variable-size batching, checkpoint dictionary,
the no-CPU-fallback contract."""
import numpy as np
import pytest
import torch

from zeroshotsemanticsegmentation_b200 import trainer as T


def items(sizes, with_vec=False, D=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for h, w in sizes:
        img = torch.randn(3, h, w, generator=g)
        lbl = torch.randint(-1, 7, (h, w), generator=g)
        out.append((img, (lbl, torch.randn(D, h, w, generator=g))) if with_vec else (img, lbl))
    return out


def test_collate_pads_images_with_zero_and_labels_with_ignore():
    its = items([(37, 53), (40, 21), (8, 64)])
    data, target = T.collate_padded(its)
    assert data.shape == (3, 3, 40, 64) and target.shape == (3, 40, 64)
    assert data.dtype == torch.float32 and target.dtype == torch.int64
    for b, (img, lbl) in enumerate(its):
        h, w = lbl.shape
        assert torch.equal(data[b, :, :h, :w], img) and torch.equal(target[b, :h, :w], lbl)
        pad = torch.ones(40, 64, dtype=torch.bool)
        pad[:h, :w] = False
        assert (target[b][pad] == T.PAD_LABEL).all() and (data[b][:, pad] == 0).all()
    # the number of valid pixels (what every loss and metric normalises by) is unchanged by padding
    assert int((target >= 0).sum()) == sum(int((l >= 0).sum()) for _, l in its)


def test_padding_stays_ignored_in_the_seenmask_target():
    """the seen-mask target maps the dataset's ignore label -1 to class 0 (upstream, trainer_seenmask.py:55-56)
    but must NOT do that to the padding collate_padded adds for B > 1, or pad regions would be trained as 'unseen'."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    a = (torch.ones(3, 4, 6), torch.tensor([[0, 1, 2, -1, 3, 4]] * 4))
    b = (torch.ones(3, 2, 3), torch.tensor([[5, -1, 2]] * 2))
    data, target = T.collate_padded([a, b])
    assert target.shape == (2, 4, 6) and T.PAD_LABEL < -1
    assert (target[1, 2:, :] == T.PAD_LABEL).all() and (target[1, :2, 3:] == T.PAD_LABEL).all()
    t = U.seenmask_target(target, unseen=[2, 5], n_class=6)
    assert t[0].tolist() == [[1, 1, 0, 0, 1, 1]] * 4                      # -1 -> 0 like upstream, unseen 2 -> 0
    assert t[1, :2, :3].tolist() == [[0, 0, 0]] * 2                        # 5 unseen, -1 -> 0, 2 unseen
    assert (t[1, 2:, :] < 0).all() and (t[1, :2, 3:] < 0).all()           # padding stays ignored
    assert (t >= 0).sum().item() == 24 + 6


def test_collate_single_item_is_the_reference_batch():
    """DataLoader(batch_size=1) of the reference (train.py:81-84) == collate_padded of one item."""
    (img, (lbl, vec)), = items([(33, 17)], with_vec=True)
    data, target = T.collate_padded([(img, (lbl, vec))])
    assert torch.equal(data, img[None]) and torch.equal(target, lbl[None].long())


def test_collate_rounding_fixed_size_and_errors():
    its = items([(37, 53), (40, 21)])
    data, target = T.collate_padded(its, multiple=32)
    assert data.shape[2:] == (64, 64)
    data, target = T.collate_padded(its, size=(48, 56))
    assert data.shape[2:] == (48, 56) and target.shape[1:] == (48, 56)
    with pytest.raises(ValueError):
        T.collate_padded(its, size=(38, 56))  # 40 rows do not fit
    with pytest.raises(ValueError):
        T.collate_padded([])
    with pytest.raises(ValueError):
        T.collate_padded([(torch.zeros(3, 4, 5), torch.zeros(4, 6))])
    # numpy items (what __getitem__ yields before the default collate) are accepted
    data, target = T.collate_padded([(np.zeros((3, 4, 5), np.float32), np.zeros((4, 5), np.int32))])
    assert data.shape == (1, 3, 4, 5) and target.dtype == torch.int64


def test_collate_works_as_dataloader_collate_fn():
    ds = items([(20, 30), (25, 18), (31, 31), (9, 40)])
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, collate_fn=T.collate_padded)
    shapes = [tuple(d.shape) for d, _ in loader]
    assert shapes == [(2, 3, 25, 30), (2, 3, 31, 40)]


class _Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(2))


def test_trainers_refuse_to_run_without_cuda():
    m = _Tiny()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    with pytest.raises(RuntimeError):
        T.Trainer(False, m, opt, [], [], None, "pascal", 1, loss_func="cross_entropy", n_class=21)
    with pytest.raises(RuntimeError):  # cuda=True but the model sits on the CPU
        T.SeenmaskTrainer(True, m, opt, [], [], None, "pascal", 1, unseen=[1], n_class=21)


def test_checkpoint_dictionary_has_the_reference_keys(tmp_path):
    m = _Tiny()
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    path = str(tmp_path / "checkpoint")
    T.save_checkpoint(path, m, opt, epoch=3, iteration=17, best_mean_iu=0.25)
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"epoch", "iteration", "arch", "optim_state_dict", "model_state_dict", "best_mean_iu"}  # trainer_fcn.py:281-288
    assert raw["arch"] == "_Tiny"
    m2 = _Tiny()
    with torch.no_grad():
        m.p.add_(1.0)
    T.save_checkpoint(path, m, opt, 4, 20, 0.5)
    ck = T.load_checkpoint(path, m2, torch.optim.SGD(m2.parameters(), lr=0.1, momentum=0.9))
    assert ck["epoch"] == 4 and ck["iteration"] == 20 and torch.equal(m2.p, m.p)
