"""End-to-end parity (GPU): FCN32s forward / backward through the drop-in module against the golden
vectors of the unmodified reference and against the CPU oracle (tf32 path: <= 1e-3 relative forward)."""
import numpy as np
import pytest
import torch

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def cos(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(a @ b / max(np.linalg.norm(a) * np.linalg.norm(b), 1e-300))


def l2(a, b):
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


# Gradient tolerances.  Every kernel is checked tightly on identical inputs in test_kernels_gpu.py (1e-3 .. 1e-6).  End to
# end, a ReLU / max-pool network is only piecewise smooth: activations that differ in the last TF32 (2^-11) or bf16
# (2^-8) bit flip a few gate / arg-max decisions, and each flip changes a gradient term by O(1).  The effect grows with
# depth (measured on the CPU oracle alone: fp32 vs TF32-storage emulation already differ by 11 % max-norm at conv1_1).
# So deep gradients are held to direction (cosine vs the fp32 oracle) and relative L2 error vs the storage-precision
# emulation, with bounds that tighten towards the head, where few decisions lie between a weight and the loss.
def grad_bounds(name, precision):
    near_head = name.startswith(("score_fr", "seenmask"))
    mid = name.startswith(("fc6", "fc7", "conv5"))
    if precision == "tf32":
        return (2e-3, 0.9999) if near_head else (6e-2, 0.995) if mid else (2.5e-1, 0.97)
    return (3e-2, 0.999) if near_head else (2e-1, 0.97) if mid else (5e-1, 0.85)


def emulated_grads(params, x, loss_fn, storage, drop_masks=None, mode="fcn"):
    """Gradients of the storage-precision emulation of the CUDA path (oracle ``storage=``): fp32 arithmetic with the
    kernels' HBM rounding points, so ReLU / max-pool decisions match the GPU and gradients compare tightly."""
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    out = O.forward(x, pr, mode, drop_masks=drop_masks, storage=storage)
    loss_fn(out).backward()
    return out, {k: v.grad for k, v in pr.items()}


def build(n_class, seed, precision="tf32"):
    import zeroshotsemanticsegmentation_b200 as szn
    m = szn.FCN32s(n_class, precision=precision)
    m.load_state_dict(O.init_params(n_class, seed))
    return m.to(DEV)


def test_forward_backward_ce21_golden(golden):
    """BASELINE configs[0] family (n_class=21, cross_entropy2d sum), small odd size 37x53."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("ce21_37x53")
    m = build(21, int(g["seed"])).eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    t = torch.from_numpy(g["target"]).long().to(DEV)
    score = m(x, mode="fcn")
    assert score.shape == (1, 21, 37, 53) and score.is_contiguous() and score.dtype == torch.float32
    e = rel(score.detach().cpu().numpy(), g["score"])
    print("forward rel err (tf32)", e)
    assert e < 1e-3
    loss = U.cross_entropy2d(score, t)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * float(g["loss"])
    loss.backward()
    # (a) against the reference's fp32 gradients: the heads are tight; deep layers differ through the ReLU / max-pool
    #     decisions that flip when activations carry TF32 (2^-11) rounding, so they are held to direction + a loose norm
    named = dict(m.named_parameters())
    for name, tol in (("score_fr.weight", 1e-2), ("score_fr.bias", 1e-2)):
        e = rel(named[name].grad.cpu().numpy(), g[name.replace(".", "__") + "__grad"])
        print(name, "grad rel err vs reference", e)
        assert e < tol
    for name in ("conv1_1.weight", "conv1_1.bias", "fc7.bias"):
        c = cos(named[name].grad.cpu().numpy(), g[name.replace(".", "__") + "__grad"])
        print(name, "grad cosine vs reference", c)
        assert c > 0.97
    assert cos(m.conv3_2.weight.grad[::16, ::16].cpu().numpy(), g["conv3_2__weight__grad_sub"]) > 0.98
    assert cos(m.fc6.weight.grad[::256, ::64].cpu().numpy(), g["fc6__weight__grad_sub"]) > 0.99
    # (b) against the storage-precision emulation of the same fp32 oracle: every gradient within 2e-3
    _, eg = emulated_grads(O.init_params(21, int(g["seed"])), torch.from_numpy(g["x"]),
                           lambda sc: O.cross_entropy2d(sc, torch.from_numpy(g["target"]).long()), "tf32")
    for name, p_ in named.items():
        if "upscore" in name or p_.grad is None or eg[name] is None:
            continue
        e = l2(p_.grad.cpu().numpy(), eg[name].numpy())
        print(name, "grad rel-L2 err vs tf32-storage oracle %.3e" % e)
        assert e < grad_bounds(name, "tf32")[0], name
    agree = (score.detach().max(1)[1].cpu().numpy() == g["lbl"]).mean()
    print("argmax agreement", agree)
    assert agree > 0.995


def test_fp32_grade_mode_matches_reference_goldens_strictly(golden):
    """precision="fp32" (3 x bf16 error-compensated tensor-core products on split hi/lo storage): forward AND every
    golden gradient of the UNMODIFIED reference, including conv1_1.weight -- the gate SURVEY 8d asks for (<= 1e-2) and the
    TF32 mode can only meet in direction -- because activations now agree to ~1e-6 and ReLU / max-pool decisions no longer
    flip."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("ce21_37x53")
    m = build(21, int(g["seed"]), precision="fp32").eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    t = torch.from_numpy(g["target"]).long().to(DEV)
    score = m(x, mode="fcn")
    e = rel(score.detach().cpu().numpy(), g["score"])
    print("forward rel err (fp32-grade)", e)
    assert e < 2e-5
    loss = U.cross_entropy2d(score, t)
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    loss.backward()
    named = dict(m.named_parameters())
    for name in ("score_fr.weight", "score_fr.bias", "fc7.bias", "conv1_1.weight", "conv1_1.bias"):
        e = rel(named[name].grad.cpu().numpy(), g[name.replace(".", "__") + "__grad"])
        print(name, "grad rel err vs reference (fp32-grade)", e)
        # SURVEY 8d's gate for conv1_1.weight is 1e-2 (a handful of ReLU gates next to zero still flip); the head is tight
        assert e < (1e-2 if name.startswith("conv1_1") else 1e-3), name
    e3 = rel(m.conv3_2.weight.grad[::16, ::16].cpu().numpy(), g["conv3_2__weight__grad_sub"])
    e6 = rel(m.fc6.weight.grad[::256, ::64].cpu().numpy(), g["fc6__weight__grad_sub"])
    print("conv3_2.weight", e3, "fc6.weight", e6)
    assert e3 < 1e-2 and e6 < 1e-3
    assert (score.detach().max(1)[1].cpu().numpy() == g["lbl"]).mean() > 0.9999
    # both heads + cosine loss on the 2 x 64 x 96 golden case
    g = golden("cos_voc20_2x64x96")
    tab = torch.from_numpy(g["table"]).float()
    m = build(tab.shape[1], int(g["seed"]), precision="fp32").eval()
    f, s = m(torch.from_numpy(g["x"]).to(DEV), mode="both")
    assert rel(f.detach().cpu().numpy(), g["score"]) < 2e-5
    assert rel(s.detach().cpu().numpy(), g["seenmask_score"]) < 2e-5
    loss = U.cosine_loss(f, torch.from_numpy(g["target"]).long().to(DEV), table=tab.to(DEV))
    loss.backward()
    assert rel(m.score_fr.weight.grad.cpu().numpy(), g["score_fr__weight__grad"]) < 1e-3
    e1 = l2(m.conv1_1.weight.grad.cpu().numpy(), g["conv1_1__weight__grad"])
    e5 = l2(m.conv5_3.weight.grad[::32, ::32].cpu().numpy(), g["conv5_3__weight__grad_sub"])
    print("rel-L2 grad error vs reference: conv1_1.weight", e1, "conv5_3.weight", e5)
    assert e1 < 1e-2 and e5 < 1e-2  # (max-norm: a single flipped ReLU gate moves one term by O(1))
    assert (U.infer_lbl(f.detach(), tab.to(DEV)) == g["lbl"]).mean() > 0.9999


def test_forward_256_golden(golden):
    """BASELINE configs[0]: 1x3x256x256, 21 classes, forward + CE loss."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("ce21_256x256")
    m = build(21, int(g["seed"])).eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    t = torch.from_numpy(g["target"]).long().to(DEV)
    with torch.no_grad():
        score = m(x)
    assert rel(score[:, :, ::8, ::8].cpu().numpy(), g["score_sub"]) < 1e-3
    loss = U.cross_entropy2d(score, t)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * float(g["loss"])


def test_embedding_heads_cos_golden(golden):
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("cos_voc20_2x64x96")
    tab = torch.from_numpy(g["table"]).float()
    m = build(tab.shape[1], int(g["seed"])).eval()
    x = torch.from_numpy(g["x"]).to(DEV)
    t = torch.from_numpy(g["target"]).long().to(DEV)
    f, s = m(x, mode="both")
    assert rel(f.detach().cpu().numpy(), g["score"]) < 1e-3
    assert rel(s.detach().cpu().numpy(), g["seenmask_score"]) < 1e-3
    loss = U.cosine_loss(f, t, table=tab.to(DEV))
    want = (g["loss_per_sample"] * g["nvalid"]).sum() / g["nvalid"].sum()
    assert abs(loss.item() - want) < 1e-4
    loss.backward()
    assert rel(m.score_fr.weight.grad.cpu().numpy(), g["score_fr__weight__grad"]) < 1e-2
    c1 = cos(m.conv1_1.weight.grad.cpu().numpy(), g["conv1_1__weight__grad"])
    c5 = cos(m.conv5_3.weight.grad[::32, ::32].cpu().numpy(), g["conv5_3__weight__grad_sub"])
    print("grad cosine vs reference: conv1_1.weight", c1, "conv5_3.weight", c5)
    assert c1 > 0.97 and c5 > 0.99
    lbl = U.infer_lbl(f.detach(), tab.to(DEV))
    print("label agreement vs reference", (lbl == g["lbl"]).mean())
    assert (lbl == g["lbl"]).mean() > 0.99


def test_mode_errors_and_seenmask_phase(golden):
    """trainer_seenmask.py:50-81: only the seenmask head is trainable; CE mean over the binary target."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    g = golden("cos_voc20_2x64x96")
    tab = torch.from_numpy(g["table"]).float()
    m = build(tab.shape[1], int(g["seed"])).eval()
    with pytest.raises(Exception, match="unexpected forward mode"):
        m(torch.zeros(1, 3, 32, 32, device=DEV), mode="nope")
    for p in m.parameters():
        p.requires_grad = False
    for p in list(m.seenmask_score.parameters()) + list(m.seenmask_upscore.parameters()):
        p.requires_grad = True
    x = torch.from_numpy(g["x"][:1]).to(DEV)
    t = torch.from_numpy(g["target"][:1]).long()
    smt = O.seenmask_target(t, list(g["train_unseen"]), tab.shape[0]).to(DEV)
    s = m(x, mode="seenmask")
    loss = U.cross_entropy2d(s, smt, size_average=True)
    assert abs(loss.item() - g["seenmask_loss"][0]) < 1e-3
    loss.backward()
    assert m.conv1_1.weight.grad is None and m.score_fr.weight.grad is None
    # oracle gradient of the head
    p = {k: v.clone().requires_grad_(k.startswith("seenmask")) for k, v in O.init_params(tab.shape[1], int(g["seed"])).items()}
    so = O.forward(torch.from_numpy(g["x"][:1]), p, "seenmask")
    O.cross_entropy2d(so, smt.cpu(), size_average=True).backward()
    assert rel(m.seenmask_score.weight.grad.cpu().numpy(), p["seenmask_score.weight"].grad.numpy()) < 1e-2
    assert rel(m.seenmask_upscore.weight.grad.cpu().numpy(), p["seenmask_upscore.weight"].grad.numpy()) < 1e-2


@pytest.mark.parametrize("precision", ["tf32", "bf16"])
def test_train_mode_dropout_masks_and_bf16(precision):
    """train(): Dropout2d masks injected so the oracle can replay them; bf16 path within its own tolerance."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    D, C, H, W, B = 20, 21, 48, 40, 2
    params = O.init_params(D, seed=21)
    x, lab, table = O.synth_batch(B, H, W, C, D, seed=21, block=8)
    g = torch.Generator().manual_seed(5)
    masks = ((torch.rand(B, 4096, generator=g) < 0.5).float(), (torch.rand(B, 4096, generator=g) < 0.5).float())
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    f_ref = O.forward(x, pr, "fcn", drop_masks=masks)
    loss_ref = O.mse_loss(f_ref, lab, O.target_embed_from_labels(lab, table))
    loss_ref.backward()
    import zeroshotsemanticsegmentation_b200 as szn
    m = szn.FCN32s(D, precision=precision)
    m.load_state_dict(params)
    m = m.to(DEV).train()
    m._forced_drop_masks = masks
    f = m(x.to(DEV))
    loss = U.mse_loss(f, lab.to(DEV), table=table.to(DEV))
    loss.backward()
    tol_f = 1e-3 if precision == "tf32" else 2e-2
    e = rel(f.detach().cpu().numpy(), f_ref.detach().numpy())
    print(precision, "forward rel err", e)
    assert e < tol_f
    # gradients: tight against the storage-precision emulation (same rounding points => same ReLU/pool decisions),
    # direction-only against the plain fp32 oracle
    f_em, eg = emulated_grads(params, x, lambda sc: O.mse_loss(sc, lab, O.target_embed_from_labels(lab, table)),
                              precision, drop_masks=masks)
    e = rel(f.detach().cpu().numpy(), f_em.detach().numpy())
    print(precision, "forward rel err vs storage-precision oracle", e)
    assert e < (5e-4 if precision == "tf32" else 4e-3)
    named = dict(m.named_parameters())
    for name, p_ in named.items():
        if "upscore" in name or name.startswith("seenmask"):
            continue
        ge = l2(p_.grad.cpu().numpy(), eg[name].numpy())
        c = cos(p_.grad.cpu().numpy(), pr[name].grad.numpy())
        print(precision, name, "grad rel-L2 err vs storage-precision oracle %.3e  cosine vs fp32 oracle %.5f" % (ge, c))
        tol_l2, tol_cos = grad_bounds(name, precision)
        assert ge < tol_l2 and c > tol_cos, name
    # random masks: about half of the channels dropped, scaled by 2
    m._forced_drop_masks = None
    f2 = m(x.to(DEV))
    assert torch.isfinite(f2).all()


def test_state_dict_and_get_parameters_contract():
    """train.py:302-331 walks named_modules(); state_dict keys/shapes must equal the reference's."""
    import torch.nn as nn
    import zeroshotsemanticsegmentation_b200 as szn
    m = szn.FCN32s(20)
    ref_shapes = {k: tuple(v.shape) for k, v in O.init_params(20).items()}
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref_shapes
    allowed = (nn.Conv2d, nn.ConvTranspose2d, nn.ReLU, nn.MaxPool2d, nn.Dropout2d, nn.Sequential, szn.FCN32s)
    for _, mod in m.named_modules():
        assert isinstance(mod, allowed)
    assert m.upscore.bias is None and m.seenmask_upscore.bias is None


def test_upscore_weight_grad_matches_the_reference_when_asked_for():
    """trainer_fcn.py:161-162 prints ``self.model.upscore.weight.grad.sum()``: autograd populates the dense (D,D,64,64)
    gradient of the never-optimised deconv (train.py:324-327).  FCN32s(upscore_weight_grad=True) reproduces it (a
    213 GFLOP/image job at D=300, hence opt-in); the default leaves ``upscore.weight.grad`` None -- the documented
    deviation: the filter is frozen, and its diagonal-bilinear structure is what makes the x32 upsample a 4-tap kernel."""
    from zeroshotsemanticsegmentation_b200 import utils as U
    import zeroshotsemanticsegmentation_b200 as szn
    D, C, H, W, B = 20, 21, 40, 56, 2
    params = O.init_params(D, seed=11)
    x, lab, table = O.synth_batch(B, H, W, C, D, seed=11, block=8)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    f_ref = O.forward(x, pr, "fcn")
    O.cosine_loss(f_ref, lab, O.target_embed_from_labels(lab, table)).backward()
    for flag in (True, False):
        m = szn.FCN32s(D, precision="fp32", upscore_weight_grad=flag)
        m.load_state_dict(params)
        m = m.to(DEV).eval()
        f = m(x.to(DEV))
        U.cosine_loss(f, lab.to(DEV), table=table.to(DEV)).backward()
        if not flag:
            assert m.upscore.weight.grad is None
            continue
        g, g_ref = m.upscore.weight.grad.cpu(), pr["upscore.weight"].grad
        assert g.shape == (D, D, 64, 64)
        e = rel(g.numpy(), g_ref.numpy())
        print("upscore.weight.grad rel err", e, " sum", float(g.sum()), "vs", float(g_ref.sum()))
        assert e < 1e-3
        assert abs(float(g.sum()) - float(g_ref.sum())) < 1e-3 * max(1.0, abs(float(g_ref.abs().sum())))
