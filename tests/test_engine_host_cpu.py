"""Host-side logic of the engine that needs no GPU: layer table, parameter order, packed-weight cache, detection of the
frozen bilinear ``upscore`` weight (models.py:102-112), optimizer argument checks."""
import pytest
import torch

from oracle import szn_oracle as O
from zeroshotsemanticsegmentation_b200 import engine, models, optim


def test_layer_table_and_parameter_order_match_the_reference_module():
    assert engine.TRUNK == O.TRUNK  # models.py:43-81
    m = models.FCN32s(7)
    names = [n for n, _ in m.named_parameters()]
    assert sorted(engine.PARAM_ORDER) == sorted(names)
    assert [p.shape for p in m._ordered_params()] == [dict(m.named_parameters())[n].shape for n in engine.PARAM_ORDER]
    # k > 1 conv weights live in channels_last memory (the wgrad kernel's output layout); values / shapes unchanged
    assert m.conv3_2.weight.is_contiguous(memory_format=torch.channels_last)
    assert m.conv1_1.weight.is_contiguous() and m.fc7.weight.shape == (4096, 4096, 1, 1)
    sd = m.state_dict()
    assert set(sd) == set(O.init_params(7)) and all(sd[k].shape == v.shape for k, v in O.init_params(7).items())
    # key order = definition order of the reference module (models.py:43-98)
    assert list(sd)[-6:] == ["score_fr.weight", "score_fr.bias", "upscore.weight", "seenmask_score.weight",
                             "seenmask_score.bias", "seenmask_upscore.weight"]
    m.load_state_dict(O.init_params(7, seed=3))   # a reference checkpoint (NCHW tensors) loads into the channels_last params
    assert m.conv3_2.weight.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(m.conv3_2.weight.detach(), O.init_params(7, seed=3)["conv3_2.weight"])


def test_packed_weight_cache_refreshes_on_version_change():
    cache = engine.PackedWeights()
    w = torch.nn.Parameter(torch.zeros(3))
    built = []

    def build():
        built.append(1)
        return len(built)

    key = lambda: (w._version, w.data_ptr())
    assert cache.get("k", key(), build) == 1 and cache.get("k", key(), build) == 1 and len(built) == 1
    with torch.no_grad():
        w.add_(1.0)                       # an optimizer step: version counter moves
    assert cache.get("k", key(), build) == 2
    w.data = torch.ones(3)                # copy_params_from_vgg16-style rebinding: data pointer moves
    assert cache.get("k", key(), build) == 3
    optim._bump_versions([w])             # what the fused optimizers do after writing parameters behind autograd's back
    assert cache.get("k", key(), build) == 4


def test_bilinear_upscore_detection():
    D = 6
    w = models.get_upsampling_weight(D, D, 64)
    assert torch.equal(w, O.upsampling_weight(D, D))
    assert engine.is_diag_bilinear(w)
    w2 = w.clone()
    w2[1, 2, 5, 5] = 1e-3                 # a trained / loaded weight that is no longer the frozen diagonal filter
    assert not engine.is_diag_bilinear(w2)
    w3 = w.clone()
    w3[0, 0, 0, 0] *= 1.5
    assert not engine.is_diag_bilinear(w3)
    assert not engine.is_diag_bilinear(torch.zeros(2, 3, 64, 64)) and not engine.is_diag_bilinear(torch.zeros(2, 2, 4, 4))
    # the filter itself: models.py:11-19 values (corner, centre, sum) of SURVEY §8a row a1
    f = models.bilinear_filter(64)
    assert abs(float(f[0, 0]) - 2.44140625e-4) < 1e-9 and abs(float(f[31, 31]) - 0.968994140625) < 1e-9
    assert abs(float(f.sum()) - 1024.0) < 1e-3


def test_forward_mode_and_precision_errors():
    m = models.FCN32s(4)
    with pytest.raises(Exception, match="unexpected forward mode"):
        m(torch.zeros(1, 3, 8, 8), mode="bogus")          # models.py:160, raised before any device work
    with pytest.raises(ValueError):
        models.FCN32s(4, precision="fp8")
    assert engine.round_up(302, 64) == 320 and engine.round_up(64, 64) == 64


def test_optimizer_argument_checks():
    p = [torch.nn.Parameter(torch.zeros(2))]
    with pytest.raises(ValueError):
        optim.FusedSGD(p, lr=-1.0)
    with pytest.raises(ValueError):
        optim.FusedAdam(p, betas=(1.0, 0.9))
    with pytest.raises(NotImplementedError):
        optim.FusedAdam(p, amsgrad=True)
    o = optim.FusedAdam([{"params": p, "lr": 2e-3}], lr=1e-3)
    assert o.param_groups[0]["lr"] == 2e-3 and o.param_groups[0]["betas"] == (0.9, 0.999)
    o.step()                                               # no gradients yet: nothing to do, nothing raised
    p[0].grad = torch.zeros(2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        o.step()
