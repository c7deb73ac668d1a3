"""One full training iteration in the reference's call order (trainer_fcn.py:83-120 forward, :149-158 train_epoch;
optimizer groups of train.py:126-129) on the CUDA path with the fused SGD step, against the CPU oracle + torch.optim.SGD."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import szn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def get_parameters(model, bias=False):
    """train.py:302-331 restated (the file itself imports tensorboardX / fcn and cannot be imported here)."""
    import zeroshotsemanticsegmentation_b200 as szn
    skipped = (nn.ReLU, nn.MaxPool2d, nn.Dropout2d, nn.Sequential, szn.FCN32s)
    for name, m in model.named_modules():
        if name in ("seenmask_score", "seenmask_upscore"):
            continue
        if isinstance(m, nn.Conv2d):
            yield m.bias if bias else m.weight
        elif isinstance(m, nn.ConvTranspose2d):
            if bias:
                assert m.bias is None
        elif isinstance(m, skipped):
            continue
        else:
            raise ValueError("Unexpected module: %s" % str(m))


def test_fused_sgd_equals_torch_sgd():
    from zeroshotsemanticsegmentation_b200.optim import FusedSGD
    g = torch.Generator().manual_seed(1)
    shapes = [(64, 32, 3, 3), (37,), (5, 7, 1, 1), (128, 64, 3, 3)]
    ref_p, my_p = [], []
    for i, s in enumerate(shapes):
        t = torch.randn(s, generator=g)
        a, b = t.clone().to(DEV), t.clone().to(DEV)
        if len(s) == 4 and s[2] > 1 and i % 2 == 0:  # channels_last parameter, like the conv weights of FCN32s
            a = a.contiguous(memory_format=torch.channels_last)
            b = b.contiguous(memory_format=torch.channels_last)
        ref_p.append(nn.Parameter(a))
        my_p.append(nn.Parameter(b))
    kw = dict(lr=1e-2, momentum=0.99, weight_decay=5e-4)
    groups = lambda ps: [{"params": ps[::2]}, {"params": ps[1::2], "lr": 2e-2, "weight_decay": 0}]
    ref, mine = torch.optim.SGD(groups(ref_p), **kw), FusedSGD(groups(my_p), **kw)
    for step in range(4):
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad = gr.clone()
            # gradients arrive either in the parameter's layout or as a strided view of another buffer
            b.grad = gr.clone() if step % 2 else gr.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2) if gr.dim() == 4 else gr.clone()
        ref.step()
        mine.step()
        for a, b in zip(ref_p, my_p):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7), step
    # state_dict layout is torch.optim.SGD's: a checkpoint of one loads into the other
    ref2 = torch.optim.SGD(groups(ref_p), **kw)
    ref2.load_state_dict(mine.state_dict())
    assert all("momentum_buffer" in s for s in ref2.state_dict()["state"].values())


def test_fused_adam_equals_torch_adam():
    """train.py:133 / :175: torch.optim.Adam(params, lr) with the bias group at lr * 2."""
    from zeroshotsemanticsegmentation_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(2)
    shapes = [(2, 4096, 1, 1), (2,), (2, 2, 64, 64), (64, 32, 3, 3), (33,)]
    ref_p, my_p = [], []
    for i, s in enumerate(shapes):
        t = torch.randn(s, generator=g)
        a, b = t.clone().to(DEV), t.clone().to(DEV)
        if s == (64, 32, 3, 3):
            a = a.contiguous(memory_format=torch.channels_last)
            b = b.contiguous(memory_format=torch.channels_last)
        ref_p.append(nn.Parameter(a))
        my_p.append(nn.Parameter(b))
    groups = lambda ps: [{"params": ps[::2]}, {"params": ps[1::2], "lr": 2e-3}]
    for kw in (dict(lr=1e-3), dict(lr=1e-3, betas=(0.8, 0.99), eps=1e-6, weight_decay=1e-2)):
        ref, mine = torch.optim.Adam(groups(ref_p), **kw), FusedAdam(groups(my_p), **kw)
        for step in range(5):
            for a, b in zip(ref_p, my_p):
                gr = torch.randn(a.shape, generator=g).to(DEV) * (10.0 ** (step - 2))
                a.grad, b.grad = gr.clone(), gr.clone()
            ref.step()
            mine.step()
            for a, b in zip(ref_p, my_p):
                assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (kw, step, (a - b).abs().max().item())
        # torch's state layout: a FusedAdam checkpoint loads into torch.optim.Adam and continues identically
        ref2 = torch.optim.Adam(groups(ref_p), **kw)
        ref2.load_state_dict(mine.state_dict())
        st = ref2.state_dict()["state"]
        assert all(set(v) >= {"step", "exp_avg", "exp_avg_sq"} and float(v["step"]) == 5 for v in st.values())
        mine2 = FusedAdam(groups(my_p), **kw)
        # (deepcopy: state_dict() hands out the live 'step' tensors and load_state_dict keeps them, as torch's own does)
        mine2.load_state_dict(copy.deepcopy(ref.state_dict()))
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, generator=g).to(DEV)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        mine2.step()
        for a, b in zip(ref_p, my_p):
            assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)
    cpu_p = nn.Parameter(torch.zeros(3))
    cpu_p.grad = torch.zeros(3)
    with pytest.raises(RuntimeError):
        FusedAdam([cpu_p], lr=1e-3).step()  # no CPU fallback


def test_training_iteration_matches_oracle():
    import zeroshotsemanticsegmentation_b200 as szn
    from zeroshotsemanticsegmentation_b200.optim import FusedSGD
    U = szn.utils
    D, C, H, W = 20, 21, 40, 56
    params = O.init_params(D, seed=9)
    x, lab, table = O.synth_batch(1, H, W, C, D, seed=9, block=8)
    lr = 1e-3
    # --- CUDA path, reference call order ---
    m = szn.FCN32s(n_class=D)
    m.load_state_dict(params)
    m = m.to(DEV).eval()  # eval: Dropout2d off so that the oracle can replay the step
    optim = FusedSGD([{"params": list(get_parameters(m, bias=False))},
                      {"params": list(get_parameters(m, bias=True)), "lr": lr * 2, "weight_decay": 0}],
                     lr=lr, momentum=0.99, weight_decay=0.0005)
    losses = []
    for it in range(2):
        score = m(x.to(DEV), mode="fcn")                                      # trainer_fcn.py:97
        loss = U.cosine_loss(score, lab.to(DEV), table=table.to(DEV))         # :100-105
        assert not np.isnan(float(loss.item()))                               # :107-108
        lbl_pred = U.infer_lbl(score, table.to(DEV))                          # :111-117
        optim.zero_grad()
        loss.backward()
        optim.step()                                                          # :156-158
        acc = U.label_accuracy_score([lab.numpy()[0]], [lbl_pred[0]], C)      # :164
        assert len(acc) == 4
        losses.append(loss.item())
    # --- oracle + torch.optim.SGD ---
    pr = {k: v.clone().requires_grad_("upscore" not in k) for k, v in params.items()}
    wts = [v for k, v in pr.items() if k.endswith("weight") and "upscore" not in k and not k.startswith("seenmask")]
    bs = [v for k, v in pr.items() if k.endswith("bias") and not k.startswith("seenmask")]
    ref_opt = torch.optim.SGD([{"params": wts}, {"params": bs, "lr": lr * 2, "weight_decay": 0}], lr=lr, momentum=0.99,
                              weight_decay=0.0005)
    ref_losses = []
    for it in range(2):
        f = O.forward(x, pr, "fcn")
        l = O.cosine_loss(f, lab, O.target_embed_from_labels(lab, table))
        ref_opt.zero_grad()
        l.backward()
        ref_opt.step()
        ref_losses.append(l.item())
    assert abs(losses[0] - ref_losses[0]) < 1e-4 and abs(losses[1] - ref_losses[1]) < 1e-3
    sd = m.state_dict()
    for name in ("score_fr.weight", "score_fr.bias", "fc7.bias", "conv5_3.weight", "conv1_1.weight"):
        a, b = sd[name].cpu(), pr[name].detach()
        upd = (b - params[name]).norm().item()
        err = (a - b).norm().item()
        print(name, "update norm %.3e  error %.3e" % (upd, err))
        assert err < (0.02 if name.startswith("score_fr") else 0.3) * upd + 1e-9
