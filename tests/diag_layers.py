"""Layer-by-layer comparison of the CUDA path with the storage-precision oracle.  NOT a test (pytest does not collect it:
no test_ prefix): a debug aid for the GPU box that lives under tests/ because only test infrastructure may import oracle/.
usage: python tests/diag_layers.py [tf32|bf16]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import szn_oracle as O
import zeroshotsemanticsegmentation_b200 as szn

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
D, C, H, W, B = 20, 21, 48, 40, 2
params = O.init_params(D, seed=21)
x, lab, table = O.synth_batch(B, H, W, C, D, seed=21, block=8)
m = szn.FCN32s(D, precision=prec); m.load_state_dict(params); m = m.cuda().eval()
f = m(x.cuda())
sv = f.grad_fn.saved
col = {}
pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
f_em = O.forward(x, pr, "fcn", collect=col, storage=prec)
col_ref = {}
O.forward(x, params, "fcn", collect=col_ref)
def rel(a, b): return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
for name in list(sv["acts"].keys()) + ["fc6", "fc7"]:
    g = sv["acts"][name] if name in sv["acts"] else sv["h6" if name == "fc6" else "h7"]
    if g is None:  # a pre-pool activation: dropped after the pool (routing codes), unless SZN_POOL_Y=1
        continue
    g = g.float().cpu().permute(0, 3, 1, 2)
    print("%-8s vs emulated %.3e   vs fp32 %.3e   (emulated vs fp32 %.3e)" % (name, rel(g, col[name].detach()), rel(g, col_ref[name]), rel(col[name].detach(), col_ref[name])))
print("score    vs emulated %.3e   vs fp32 %.3e" % (rel(f.detach().cpu(), f_em.detach()), rel(f.detach().cpu(), O.forward(x, params, "fcn"))))
# backward
loss = szn.utils.mse_loss(f, lab.cuda(), table=table.cuda()); loss.backward()
O.mse_loss(f_em, lab, O.target_embed_from_labels(lab, table)).backward()
for n, p_ in m.named_parameters():
    if p_.grad is not None and "upscore" not in n:
        print("%-24s grad vs emulated %.3e" % (n, rel(p_.grad.cpu(), pr[n].grad)))
