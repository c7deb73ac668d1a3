"""2-GPU debug of bench.py's grad_check (torchrun --nproc-per-node 2 tools/debug_ddp.py [fused])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import zeroshotsemanticsegmentation_b200 as szn
from zeroshotsemanticsegmentation_b200 import ddp, synth
U = szn.utils
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D, C, H = (int(os.environ.get("DBG_B", "8")), 300, 59, int(os.environ.get("DBG_H", "512")))
def build():
    return synth.init_model_(szn.FCN32s(D, precision="tf32", fused_head=fused), seed=1337).to(dev).eval()
def batch(r):
    return synth.synth_batch(B, H, H, C, D, seed=1337 + r)
model = build()
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
red = ddp.GradientAllReduce(model)
chg = max(float((v - sd0[k]).abs().max()) for k, v in model.state_dict().items())
x, lab, _ = batch(rank); table = batch(0)[2]; x, lab, table = x.to(dev), lab.to(dev), table.to(dev)
accs = []
def hook(a):
    accs.append(a.clone()); red.accum_hook(a); accs.append(a.clone())
model.zero_grad(set_to_none=True)
f = model(x, mode="fcn")
loss = U.cosine_loss(f, lab, table=table, accum_hook=hook)
loss.backward()
torch.cuda.synchronize()
print("rank", rank, "param change by broadcast", chg, "local accum", accs[0].tolist(), "global", accs[1].tolist(), "loss", loss.item(), "f sum", float(f.double().sum()), flush=True)
dist.barrier()
if rank == 0:
    ref = build(); ref.load_state_dict(model.state_dict())
    tot = []
    shards = [(x, lab)] + [tuple(t.to(dev) for t in batch(r)[:2]) for r in range(1, world)]
    with torch.no_grad():
        for xs, ls in shards:
            ff = ref(xs, mode="fcn")
            U.cosine_loss(ff, ls, table=table, accum_hook=lambda a: tot.append(a.clone()))
            print("replica shard accum", tot[-1].tolist(), "f sum", float(ff.double().sum()), flush=True)
    total = sum(tot)
    for xs, ls in shards:
        l1 = U.cosine_loss(ref(xs, mode="fcn"), ls, table=table, accum_hook=lambda a: a.copy_(total))
        l1.backward()
    print("replica loss", l1.item())
    g1 = dict(ref.named_parameters())
    for n, p in model.named_parameters():
        if p.grad is not None:
            e = float((p.grad - g1[n].grad).norm() / g1[n].grad.norm())
            if e > 1e-5 or n in ("score_fr.weight", "conv1_1.weight"):
                print("  %-22s rel-L2 %.3e  |g| %.3e vs %.3e" % (n, e, float(p.grad.norm()), float(g1[n].grad.norm())))
dist.barrier(); dist.destroy_process_group()
