"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

    python oracle/gen_golden.py

Weights are not stored (134 M values): every case rebuilds them with
``szn_oracle.init_params(n_class, seed)`` (a seeded torch CPU generator, reproducible on the GPU box,
same image) and loads them into the reference ``models.FCN32s`` with ``load_state_dict``; a checksum
of the parameters is stored so drift is detected.  Inputs, labels and tables are stored.
The reference's loss/inference functions only work for n == 1: batched cases loop the
reference over samples and store per-sample results.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, szn_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def param_checksum(p):
    return float(sum(v.double().abs().sum() for k, v in sorted(p.items()) if "upscore" not in k))


def ref_model(M, n_class, seed):
    p = O.init_params(n_class, seed)
    m = M.FCN32s(n_class)
    m.load_state_dict(p, strict=True)
    return m, p


def grads_of(m, names):
    return {n.replace(".", "__") + "__grad": dict(m.named_parameters())[n].grad.numpy().copy() for n in names}


def case_ce(M, U, name, H, W, seed, store_score=True):
    """BASELINE config 1 shape family: n_class=21, cross_entropy2d(sum) (trainer_fcn.py:97-105)."""
    m, p = ref_model(M, 21, seed)
    m.eval()
    x, lab, _ = O.synth_batch(1, H, W, 21, 21, seed=seed, block=8)
    score = m(x, mode="fcn")
    loss = U.cross_entropy2d(score, lab, size_average=False)
    loss.backward()
    d = dict(x=x.numpy(), target=lab.numpy(), loss=np.float64(loss.item()), seed=seed, n_class=21,
             param_checksum=param_checksum(p), score_sum=np.float64(score.double().sum().item()),
             lbl=score.detach().max(1)[1].numpy().astype(np.int16),
             **grads_of(m, ["score_fr.weight", "conv1_1.weight", "conv1_1.bias", "score_fr.bias"]))
    d["conv3_2__weight__grad_sub"] = m.conv3_2.weight.grad[::16, ::16].numpy().copy()
    d["fc6__weight__grad_sub"] = m.fc6.weight.grad[::256, ::64].numpy().copy()
    d["fc7__bias__grad"] = m.fc7.bias.grad.numpy().copy()
    if store_score:
        d["score"] = score.detach().numpy()
    else:
        d["score_sub"] = score.detach()[:, :, ::8, ::8].numpy().copy()
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "loss", loss.item())


def case_embed(M, U, name, B, H, W, table, loss_name, seed, unseen=(), train_unseen=()):
    """Embedding head (trainer_fcn.py:83-147): cosine/mse loss, infer_lbl, SZN stitch; seenmask CE-mean
    (trainer_seenmask.py:50-70).  Reference looped per sample."""
    C, D = table.shape
    m, p = ref_model(M, D, seed)
    m.eval()
    x, lab, _ = O.synth_batch(B, H, W, C, D, seed=seed, block=8)
    tab = torch.from_numpy(table).float()
    te = O.target_embed_from_labels(lab, tab)
    seen_tab, unseen_tab = O.split_tables(tab, list(unseen))
    losses, nvalid, lbls, lbls_szn, lbls_forced, sm_losses = [], [], [], [], [], []
    fn = {"cos": U.cosine_loss, "mse": U.mse_loss}[loss_name]
    scores, sscores = [], []
    for i in range(B):
        f, s = m(x[i:i + 1], mode="both")
        loss = fn(f, lab[i:i + 1], te[i:i + 1])
        losses.append(loss.item())
        nvalid.append(int((lab[i] >= 0).sum()))
        lbls.append(U.infer_lbl(f, tab))
        if unseen:
            lbls_szn.append(U.infer_lbl_szn(f, s, seen_tab, unseen_tab))
            lbls_forced.append(U.infer_lbl_forced_unseen(f, lab[i:i + 1], seen_tab, unseen_tab, list(unseen)))
            smt = O.seenmask_target(lab[i:i + 1], list(train_unseen), C)
            sm_losses.append(U.cross_entropy2d(s, smt, size_average=True).item())
        scores.append(f.detach().numpy())
        sscores.append(s.detach().numpy())
        # batched-loss gradient = sum_i N_i/N * grad_i  (the loss is normalised by the global count)
        (loss * nvalid[-1]).backward()
    ntot = sum(nvalid)
    g = {k: v / ntot for k, v in grads_of(m, ["score_fr.weight", "score_fr.bias", "conv1_1.weight"]).items()}
    g["conv5_3__weight__grad_sub"] = (m.conv5_3.weight.grad[::32, ::32] / ntot).numpy().copy()
    d = dict(x=x.numpy(), target=lab.numpy().astype(np.int16), table=table, loss_per_sample=np.array(losses),
             nvalid=np.array(nvalid), lbl=np.concatenate(lbls).astype(np.int16), seed=seed,
             param_checksum=param_checksum(p), unseen=np.array(list(unseen), dtype=np.int64),
             train_unseen=np.array(list(train_unseen), dtype=np.int64),
             score=np.concatenate(scores), seenmask_score=np.concatenate(sscores), **g)
    if unseen:
        d.update(lbl_szn=np.concatenate(lbls_szn).astype(np.int16),
                 lbl_forced=np.concatenate(lbls_forced).astype(np.int16), seenmask_loss=np.array(sm_losses))
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "losses", losses)


def case_head_only(U, name, table, h, w, seed, unseen):
    """Loss / inference functions alone on an arbitrary score tensor (utils.py:19-102,159-205), incl. the
    edge cases: ignore label -1, zero rows that win when every live cosine is negative, exact ties."""
    g = torch.Generator().manual_seed(seed)
    C, D = table.shape
    tab = torch.from_numpy(table).float()
    score = torch.randn(1, D, h, w, generator=g)
    # rows 0..1 of the image: score = -(sum of all class vectors) -> every live cosine negative-ish
    score[0, :, 0, :] = -tab.sum(0)[:, None]
    # row 2: score equals class 3's vector exactly; row 3: the exact midpoint of classes 1 and 2 (near tie)
    score[0, :, 2, :] = tab[3][:, None]
    lab = torch.randint(-1, C, (1, h, w), generator=g)
    te = O.target_embed_from_labels(lab, tab)
    seen_tab, unseen_tab = O.split_tables(tab, list(unseen))
    sm = torch.randn(1, 2, h, w, generator=g)
    sc = score.clone().requires_grad_(True)
    lc = U.cosine_loss(sc, lab, te); lc.backward(); gcos = sc.grad.numpy().copy()
    sc = score.clone().requires_grad_(True)
    lm = U.mse_loss(sc, lab, te); lm.backward(); gmse = sc.grad.numpy().copy()
    ce_in = torch.randn(1, 21, h, w, generator=g)
    ce_lab = torch.randint(-1, 21, (1, h, w), generator=g)
    sc = ce_in.clone().requires_grad_(True)
    lce = U.cross_entropy2d(sc, ce_lab); lce.backward(); gce = sc.grad.numpy().copy()
    sc2 = sm.clone().requires_grad_(True)
    smt = O.seenmask_target(lab, list(unseen), C)
    lsm = U.cross_entropy2d(sc2, smt, size_average=True); lsm.backward(); gsm = sc2.grad.numpy().copy()
    d = dict(score=score.numpy(), target=lab.numpy().astype(np.int16), table=table, seenmask_score=sm.numpy(),
             unseen=np.array(list(unseen)), cos_loss=lc.item(), mse_loss=lm.item(), cos_grad=gcos, mse_grad=gmse,
             ce_score=ce_in.numpy(), ce_target=ce_lab.numpy().astype(np.int16), ce_loss=lce.item(), ce_grad=gce,
             sm_loss=lsm.item(), sm_grad=gsm,
             lbl=U.infer_lbl(score, tab).astype(np.int16),
             lbl_seen_only=U.infer_lbl(score, seen_tab).astype(np.int16),
             lbl_szn=U.infer_lbl_szn(score, sm, seen_tab, unseen_tab).astype(np.int16),
             lbl_forced=U.infer_lbl_forced_unseen(score, lab, seen_tab, unseen_tab, list(unseen)).astype(np.int16))
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, lc.item(), lm.item(), lce.item(), lsm.item())


def main():
    os.makedirs(OUT, exist_ok=True)
    M, U = ref_import.load_reference()
    root = ref_import.reference_root()
    ctx300 = U.load_obj(os.path.join(root, "datasets/context/embeddings/norm_embed_arr_300"))
    voc20 = U.load_obj(os.path.join(root, "datasets/pascal/embeddings/norm_embed_arr_20"))
    np.savez_compressed(os.path.join(OUT, "tables"), context300=ctx300, pascal20=voc20)
    case_ce(M, U, "ce21_37x53", 37, 53, 1337)
    case_ce(M, U, "ce21_256x256", 256, 256, 1338, store_score=False)     # BASELINE configs[0]
    case_embed(M, U, "cos_voc20_2x64x96", 2, 64, 96, voc20, "cos", 1339, unseen=(5, 9, 17), train_unseen=(5,))
    case_embed(M, U, "mse_voc20_1x45x70", 1, 45, 70, voc20, "mse", 1340)
    case_head_only(U, "head_ctx300_24x40", ctx300, 24, 40, 1341, unseen=(2, 30))
    case_head_only(U, "head_voc20_33x17", voc20, 33, 17, 1342, unseen=(0, 12))


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
