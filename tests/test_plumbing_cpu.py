"""The Python glue of the GPU path, exercised on the CPU with the kernels stubbed out: every C-ABI call made by
``engine.FCN32sFunction`` (forward and backward), the losses, the label functions and the experimental fused head is
checked against the ctypes signature it is bound with (argument count, integer vs pointer), and autograd checks every
returned gradient against its parameter's shape.  Values are garbage (no kernel runs): this pins control flow, shapes and
call order only — numerics are the GPU tests' job."""
import ctypes

import pytest
import torch

from zeroshotsemanticsegmentation_b200 import _lib, engine, models, utils


class Recorder:
    def __init__(self):
        self.names = []
        self.args = []

    def __call__(self, name, *args):
        sig = _lib.SIGNATURES[name]
        assert len(args) == len(sig), "%s: %d arguments for %d parameters" % (name, len(args), len(sig))
        for i, (a, t) in enumerate(zip(args, sig)):
            if t is ctypes.c_void_p:
                assert a is None or isinstance(a, int), "%s arg %d: pointer expected, got %r" % (name, i, type(a))
            elif t is ctypes.c_float:
                assert isinstance(a, float), "%s arg %d: float expected, got %r" % (name, i, type(a))
            else:
                assert isinstance(a, int) and not isinstance(a, bool), "%s arg %d: int expected, got %r" % (name, i, type(a))
        self.names.append(name)
        self.args.append(args)


@pytest.fixture()
def stubbed(monkeypatch):
    rec = Recorder()
    for mod in (_lib, engine, utils):
        monkeypatch.setattr(mod, "call", rec, raising=True)
    monkeypatch.setattr(_lib, "stream", lambda: 0)
    monkeypatch.setattr(utils, "_check_cuda", lambda *ts: None)
    return rec


def run(model, x, mode="fcn"):
    out = engine.FCN32sFunction.apply(model, x, *model._ordered_params())
    if len(out) == 3:
        out[0]._szn_head = engine.ScoreHandle(out[2], out[0])
    return out[0], out[1]


D, C, H, W = 5, 7, 8, 12
POOL_FWD, POOL_BWD = ("szn_pool_fwd", "szn_pool_bwd") if engine._POOL_Y else ("szn_pool_fwd_code", "szn_pool_bwd_code")


def inputs(B=1):
    g = torch.Generator().manual_seed(0)
    return (torch.randn(B, 3, H, W, generator=g), torch.randint(-1, C, (B, H, W), generator=g),
            torch.randn(C, D, generator=g))


def test_pool_routing_codes_replace_the_pre_pool_activations(stubbed, monkeypatch):
    """Default: szn_pool_fwd_code writes a (B, Ho, Wo, C) uint8 code tensor per pool and the pre-pool activation is dropped
    from the saved state; SZN_POOL_Y=1 (engine._POOL_Y) keeps the activation and the y-reading pair."""
    monkeypatch.setattr(engine, "_POOL_Y", False)
    m = models.FCN32s(D).train()
    x, lab, table = inputs(1)
    f, s = run(m, x)
    sv = f.grad_fn.saved
    assert sorted(sv["codes"]) == ["pool1", "pool2", "pool3", "pool4", "pool5"]
    for pool, conv in (("pool1", "conv1_2"), ("pool2", "conv2_2"), ("pool3", "conv3_3"), ("pool4", "conv4_3"), ("pool5", "conv5_3")):
        h, w, c = sv["dims"][pool]
        assert sv["codes"][pool].shape == (1, h, w, c) and sv["codes"][pool].dtype == torch.uint8
        assert sv["acts"][conv] is None and sv["acts"][pool] is not None
    i = stubbed.names.index("szn_pool_fwd_code")
    a = stubbed.args[i]
    assert a[3] == sv["codes"]["pool1"].data_ptr() and tuple(a[4:8]) == (1,) + sv["dims"]["conv1_2"]
    utils.cosine_loss(f, lab, table=table).backward()
    i = stubbed.names.index("szn_pool_bwd_code")     # the first pool reached backwards is pool5
    a = stubbed.args[i]
    assert a[1] == sv["codes"]["pool5"].data_ptr() and tuple(a[4:8]) == (1,) + sv["dims"]["conv5_3"] and a[8] == 1
    del stubbed.names[:], stubbed.args[:]
    monkeypatch.setattr(utils, "_check_cuda", lambda *ts: None)
    for frozen in (False, True):                     # inference / frozen trunk: no pool backward will run, plain pools
        monkeypatch.setattr(m, "_grad_enabled", frozen, raising=False)
        for n, p in m.named_parameters():
            p.requires_grad_(not (frozen and n.startswith("conv")))
        run(m, x)
        assert stubbed.names.count("szn_pool_fwd") == 5 and "szn_pool_fwd_code" not in stubbed.names
        del stubbed.names[:], stubbed.args[:]
    for p in m.parameters():
        p.requires_grad_(True)
    monkeypatch.setattr(m, "_grad_enabled", True, raising=False)
    run(m, x)
    assert stubbed.names.count("szn_pool_fwd_code") == 5
    monkeypatch.setattr(engine, "_POOL_Y", True)
    del stubbed.names[:], stubbed.args[:]
    f, s = run(m, x)
    assert stubbed.names.count("szn_pool_fwd") == 5 and "szn_pool_fwd_code" not in stubbed.names
    assert f.grad_fn.saved["codes"] == {} and f.grad_fn.saved["acts"]["conv1_2"] is not None
    utils.cosine_loss(f, lab, table=table).backward()
    assert stubbed.names.count("szn_pool_bwd") == 5 and "szn_pool_bwd_code" not in stubbed.names


def test_default_path_calls_and_gradient_shapes(stubbed):
    m = models.FCN32s(D).train()
    x, lab, table = inputs(2)
    f, s = run(m, x)
    assert f.shape == (2, D, H, W) and s.shape == (2, 2, H, W) and f.is_contiguous()
    fwd = list(stubbed.names)
    assert fwd[0] == "szn_conv1_1_fwd" and fwd.count("szn_conv_fwd") == 15 and fwd.count(POOL_FWD) == 5
    assert fwd[-2:] == ["szn_upsample32_crop_fwd", "szn_deconv_small_fwd"] and "szn_dropout_scale" in fwd
    loss = utils.cosine_loss(f, lab, table=table)
    del stubbed.names[:]
    loss.backward()   # autograd validates every gradient's shape against its parameter
    bwd = stubbed.names
    assert bwd[0] == "szn_embed_loss_bwd" and "szn_upsample32_crop_bwd" in bwd
    assert bwd.count("szn_conv_wgrad") == 15 and bwd.count("szn_conv_dgrad") == 15 and bwd.count(POOL_BWD) == 5
    assert bwd[-1] == "szn_conv1_1_wgrad"
    for n, p in m.named_parameters():
        if "upscore" in n or n.startswith("seenmask"):
            assert p.grad is None, n                     # unused head / frozen filter: None, not zeros
        else:
            assert p.grad is not None and p.grad.shape == p.shape, n
    assert m.conv3_2.weight.grad.stride() == m.conv3_2.weight.stride()   # the wgrad buffer is adopted without a copy
    labels = utils.infer_lbl_device(f.detach(), table)
    assert labels.shape == (2, H, W) and labels.dtype == torch.int64


def test_both_heads_and_frozen_trunk(stubbed):
    m = models.FCN32s(D).eval()
    x, lab, table = inputs()
    for p in m.parameters():
        p.requires_grad = False
    for p in list(m.seenmask_score.parameters()) + list(m.seenmask_upscore.parameters()):
        p.requires_grad = True
    f, s = run(m, x)
    assert "szn_dropout_scale" not in stubbed.names      # eval mode: no masks
    loss = utils.cross_entropy2d(s, (lab >= 0).long(), size_average=True, accum_hook=lambda acc: None)
    assert stubbed.names[-1] == "szn_loss_finalize"      # the hook re-finalises the loss from the (all-reduced) accumulator
    del stubbed.names[:]
    loss.backward()
    assert "szn_conv_dgrad" not in stubbed.names and stubbed.names.count("szn_conv_wgrad") == 1   # head GEMM only
    assert m.seenmask_score.weight.grad.shape == (2, 4096, 1, 1) and m.seenmask_upscore.weight.grad.shape == (2, 2, 64, 64)
    assert m.score_fr.weight.grad is None and m.conv1_1.weight.grad is None
    out = utils._stitch(f.detach(), table, table, seen_mask_score=s.detach())
    assert out.shape == (1, H, W)


def test_experimental_fused_head_routes_through_the_score_map(stubbed):
    m = models.FCN32s(D, fused_head=True).eval()
    x, lab, table = inputs(2)
    f, s = run(m, x)
    head = utils._fused_handle(f)
    assert head is not None and head.s17.shape == (2, 1, 1, 64) and head.D == D and head.hw == (H, W)
    assert utils._fused_handle(f.detach()) is None       # another tensor object: ordinary path
    del stubbed.names[:]
    loss = utils.cosine_loss(f, lab, table=table)
    assert stubbed.names == ["szn_head_fused_fwd"]
    labels = utils.infer_lbl_device(f, table)            # the loss pass already wrote them: no second launch
    assert stubbed.names == ["szn_head_fused_fwd"] and labels.shape == (2, H, W) and labels is head.labels_cache[3]
    other = table.clone()
    assert utils.infer_lbl_device(f, other).shape == (2, H, W)   # another table: its own label-only launch
    assert stubbed.names == ["szn_head_fused_fwd"] * 2
    utils.mse_loss(f, lab, table=table)
    assert stubbed.names == ["szn_head_fused_fwd"] * 3
    del stubbed.names[:]
    loss.backward()
    assert stubbed.names[0] == "szn_head_fused_bwd" and stubbed.names[1] == "szn_cast"
    assert "szn_upsample32_crop_bwd" not in stubbed.names and "szn_embed_loss_bwd" not in stubbed.names
    assert m.score_fr.weight.grad.shape == m.score_fr.weight.shape and m.conv1_1.weight.grad is not None
    assert m.seenmask_score.weight.grad is None
    # an explicit target_embed, or a score that was modified, takes the ordinary path
    f2, _ = run(m, x)
    del stubbed.names[:]
    utils.cosine_loss(f2, lab, torch.zeros(2, D, H, W))
    assert stubbed.names == ["szn_embed_loss_fwd"]
    f3, _ = run(m, x)
    with torch.no_grad():
        f3.mul_(2.0)
    assert utils._fused_handle(f3) is None


def test_optimizers_call_their_kernels_with_bound_signatures(stubbed, monkeypatch):
    from zeroshotsemanticsegmentation_b200 import optim
    monkeypatch.setattr(optim, "call", stubbed)
    p = torch.nn.Parameter(torch.zeros(4, 3, 3, 3).contiguous(memory_format=torch.channels_last))
    p.grad = torch.zeros(4, 3, 3, 3)                      # NCHW gradient for a channels_last parameter: re-laid out
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))  # let the CPU tensors pass the device check
    for opt in (optim.FusedSGD([p], lr=0.1, momentum=0.9), optim.FusedAdam([p], lr=1e-3)):
        v0 = p._version
        opt.step()
        assert p._version == v0 + 1                       # the packed-weight cache sees the update
    assert stubbed.names[-2:] == ["szn_sgd_step", "szn_adam_step"]


class _Loader(list):
    def __init__(self, batches, n_class=C):
        super().__init__(batches)
        import types
        self.dataset = types.SimpleNamespace(class_names=["c%d" % i for i in range(n_class)])


@pytest.fixture()
def trainer_stubs(stubbed, monkeypatch):
    from zeroshotsemanticsegmentation_b200 import optim, trainer
    monkeypatch.setattr(optim, "call", stubbed)
    monkeypatch.setattr(trainer, "_require_cuda", lambda device: None)
    monkeypatch.setattr(trainer._Base, "_check_loss", staticmethod(lambda loss: 0.0))   # the stubbed loss is uninitialised memory
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))            # CPU tensors pass the device checks
    return stubbed


@pytest.mark.parametrize("loss_func,forced", [("cos", False), ("mse", True), ("cross_entropy", False)])
def test_fcn_trainer_control_flow(trainer_stubs, tmp_path, loss_func, forced):
    """trainer_fcn.py:83-158,181-292 through trainer.Trainer with stubbed kernels: all three losses, forced-unseen and
    stitched inference, logs and checkpoint; the reference-style (lbl, lbl_vec) item and the labels-only item."""
    import os
    from zeroshotsemanticsegmentation_b200 import optim, trainer
    emb = loss_func != "cross_entropy"
    m = models.FCN32s(D if emb else C)
    x, lab, table = inputs(2)
    vec = table[lab.clamp(min=0)].permute(0, 3, 1, 2).contiguous()
    batches = [(x, lab), (x, (lab, vec))] if emb else [(x, lab)]
    ps = [p for n, p in m.named_parameters() if "upscore" not in n and not n.startswith("seenmask")]
    tr = trainer.Trainer(True, m, optim.FusedSGD(ps, lr=1e-3, momentum=0.99), _Loader(batches), _Loader(batches),
                         str(tmp_path), "pascal", 1, None, pixel_embeddings=D if emb else None, loss_func=loss_func,
                         unseen=[1, 4], val_unseen=[4], forced_unseen=forced, embed_arr=table.numpy() if emb else None)
    tr.verbose = False
    tr.train()
    assert tr.iteration == len(batches)
    names = trainer_stubs.names
    want_loss = "szn_ce2d_fwd" if loss_func == "cross_entropy" else "szn_embed_loss_fwd"
    assert want_loss in names and "szn_sgd_step" in names and "szn_confusion_hist" in names
    assert ("szn_stitch_labels" in names) == (emb and forced)
    assert ("szn_embed_argmax" in names) == emb
    assert os.path.isfile(os.path.join(str(tmp_path), "checkpoint")) and os.path.isfile(os.path.join(str(tmp_path), "val_log.csv"))
    if emb:
        del names[:]
        tr.validate(both_fcn_and_seenmask=True)
        assert "szn_stitch_labels" in names
        score, loss, lbl_pred, lbl_true = tr.forward_szn(x, (lab, vec))
        assert lbl_pred.shape == (2, H, W) and lbl_true.shape == (2, H, W)
    else:
        with pytest.raises(ValueError):
            tr.forward_szn(x, lab)
    with pytest.raises(ValueError):
        trainer.Trainer(True, m, None, _Loader([]), _Loader([]), None, "pascal", 1, loss_func="bogus", n_class=C)


def test_seenmask_trainer_control_flow(trainer_stubs, tmp_path):
    from zeroshotsemanticsegmentation_b200 import optim, trainer
    m = models.FCN32s(D)
    x, lab, table = inputs(1)
    head = trainer.freeze_for_seenmask(m)
    tr = trainer.SeenmaskTrainer(True, m, optim.FusedAdam([{"params": head}], lr=1e-3), _Loader([(x, lab)]),
                                 _Loader([(x, (lab, None))]), str(tmp_path), "pascal", 1, None, checkpoint={"epoch": 3},
                                 unseen=[1, 4])
    tr.verbose = False
    tr.train()
    names = trainer_stubs.names
    assert names.count("szn_adam_step") == 3 and "szn_conv_dgrad" not in names and "szn_ce2d_bwd" in names
    best = torch.load(str(tmp_path / "best"), weights_only=False)
    assert best["epoch"] == 3 and "seenmask_score.weight" in best["model_state_dict"]


def test_rare_engine_routes(stubbed, monkeypatch):
    """Paths the benchmark never takes: the dense upscore weight gradient the reference computes and discards
    (trainer_fcn.py:161-162), a trained (non-bilinear) upscore weight, bf16 storage, the data-parallel gradient hooks."""
    x, lab, table = inputs(1)
    # upscore_weight_grad=True: .grad of the (D,D,64,64) deconv weight is produced
    m = models.FCN32s(D, upscore_weight_grad=True).eval()
    f, s = run(m, x)
    utils.mse_loss(f, lab, table=table).backward()
    assert m.upscore.weight.grad is not None and m.upscore.weight.grad.shape == (D, D, 64, 64)
    assert "szn_deconv_small_wgrad" in stubbed.names
    # a trained upscore weight is no longer the frozen diagonal filter: dense small-deconv kernels on both passes
    m = models.FCN32s(D).eval()
    with torch.no_grad():
        m.upscore.weight[0, 1, 3, 3] = 0.5
    del stubbed.names[:]
    f, s = run(m, x)
    assert stubbed.names.count("szn_deconv_small_fwd") == 2 and "szn_upsample32_crop_fwd" not in stubbed.names
    utils.cosine_loss(f, lab, table=table).backward()
    assert "szn_deconv_small_dgrad" in stubbed.names and "szn_upsample32_crop_bwd" not in stubbed.names
    # bf16 storage: activations are bf16 tensors, the public outputs stay fp32
    m = models.FCN32s(D, precision="bf16").eval()
    f, s = run(m, x)
    assert f.dtype == torch.float32 and s.dtype == torch.float32
    (f.sum() + s.sum()).backward()
    assert m.conv2_1.weight.grad.dtype == torch.float32 and m.seenmask_upscore.weight.grad.shape == (2, 2, 64, 64)
    # data-parallel hooks: every parameter gradient is announced exactly once, the flush comes last
    m = models.FCN32s(D).eval()
    seen, order = [], []
    m._grad_ready = lambda name, g: (seen.append(name), order.append("g"))
    m._grad_flush = lambda: order.append("flush")
    f, s = run(m, x)
    utils.cosine_loss(f, lab, table=table).backward()
    expected = {n for n, p in m.named_parameters() if "upscore" not in n and not n.startswith("seenmask")}
    assert sorted(seen) == sorted(expected) and order[-1] == "flush" and order.count("flush") == 1
    assert seen[0].startswith("score_fr") and seen.index("fc6.weight") < seen.index("conv1_1.weight")  # fc6's 411 MB go first


def test_inference_helpers_return_numpy_like_the_reference(stubbed):
    x, lab, table = inputs(2)
    score = torch.randn(2, D, H, W)
    sm = torch.randn(2, 2, H, W)
    seen_t, unseen_t = utils.split_embeddings(table, [1, 4])
    import numpy as np
    for out in (utils.infer_lbl(score, table), utils.infer_lbl_szn(score, sm, seen_t, unseen_t),
                utils.infer_lbl_forced_unseen(score, lab, seen_t, unseen_t, [1, 4]),
                utils.stich_seen_unseen_with_mask(score, seen_t, unseen_t, np.zeros((2, H, W), dtype=bool))):
        assert isinstance(out, np.ndarray) and out.dtype == np.int64 and out.shape == (2, H, W)
    assert stubbed.names.count("szn_stitch_labels") == 2 and stubbed.names.count("szn_embed_argmax") == 7
    with pytest.raises(ValueError):
        utils.infer_lbl(score, torch.zeros(C, D + 1))
    with pytest.raises(NotImplementedError):
        utils.cross_entropy2d(score, lab, weight=torch.ones(D))


def test_engine_rejects_what_the_kernels_cannot_read(stubbed):
    m = models.FCN32s(D).eval()
    x, _, _ = inputs(1)
    with pytest.raises(TypeError):
        run(m, x.double())
    with pytest.raises(ValueError):
        run(m, torch.zeros(1, 4, H, W))
    with pytest.raises(NotImplementedError):
        run(m, x.clone().requires_grad_(True))
    m.fc7.double()
    with pytest.raises(TypeError, match="fc7.weight"):
        run(m, x)
    assert stubbed.names == []   # nothing was launched


def test_losses_and_labels_reject_mismatched_shapes(stubbed):
    """The kernels index raw memory; a wrong shape must raise before anything is launched (the reference raises inside
    torch for the same mistakes)."""
    score = torch.randn(2, D, H, W)
    lab = torch.zeros(2, H, W, dtype=torch.long)
    table = torch.randn(C, D)
    bad = [lambda: utils.cosine_loss(score, lab[:1], table=table),
           lambda: utils.cosine_loss(score, lab, torch.zeros(2, D, H, W + 1)),
           lambda: utils.mse_loss(score, lab, table=torch.randn(C, D + 1)),
           lambda: utils.mse_loss(score, lab),
           lambda: utils.cross_entropy2d(score, lab[:, :-1]),
           lambda: utils.cross_entropy2d(score[0], lab),
           lambda: utils.infer_lbl(score, torch.randn(C, D - 1)),
           lambda: utils.infer_lbl_szn(score, torch.zeros(2, 3, H, W), table, table),
           lambda: utils.infer_lbl_forced_unseen(score, lab[:1], table, table, [1]),
           lambda: utils.confusion_hist_device(lab, lab[:1], C)]
    for fn in bad:
        with pytest.raises(ValueError):
            fn()
    assert "szn_embed_loss_fwd" not in stubbed.names and "szn_ce2d_fwd" not in stubbed.names
    assert "szn_stitch_labels" not in stubbed.names and "szn_confusion_hist" not in stubbed.names
