"""The "library GPU" bar of SURVEY §8d: the SAME network and step (forward + cosine loss + backward + nearest-embedding
labels, train mode) written with stock torch.nn modules, i.e. cuDNN / cuBLAS kernels chosen by PyTorch, timed on the B200
next to our path.  The reference ships no GPU kernels of its own, so this is what "the reference on a B200" means.

    python tools/library_bar.py [--D 300] [--C 59] [--B 8] [--steps 5] [--upscore dense|grouped] [--dtype tf32|bf16]

--upscore dense   : ConvTranspose2d(D, D, 64, stride=32) as written (models.py:94): 213 GF/image wasted at D=300
--upscore grouped : the same frozen bilinear filter as a depthwise (groups=D) transposed conv: identical output, the most
                    favourable way to run the reference's upsampling through the library
Not part of the product or of any parity claim; plain torch only (no oracle import, no libszn).  Prints one JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F

from zeroshotsemanticsegmentation_b200 import synth
from zeroshotsemanticsegmentation_b200.models import bilinear_filter

CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]


class StockFCN32s(nn.Module):
    def __init__(self, D, upscore):
        super().__init__()
        layers, cin, first = [], 3, True
        for v in CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, stride=2, ceil_mode=True))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=100 if first else 1), nn.ReLU(inplace=True)]
                cin, first = v, False
        layers += [nn.Conv2d(512, 4096, 7), nn.ReLU(inplace=True), nn.Dropout2d(),
                   nn.Conv2d(4096, 4096, 1), nn.ReLU(inplace=True), nn.Dropout2d()]
        self.trunk = nn.Sequential(*layers)
        self.score_fr = nn.Conv2d(4096, D, 1)
        self.seenmask_score = nn.Conv2d(4096, 2, 1)
        self.grouped = upscore == "grouped"
        self.upscore = nn.ConvTranspose2d(D, D, 64, stride=32, bias=False, groups=D if self.grouped else 1)
        self.seenmask_upscore = nn.ConvTranspose2d(2, 2, 64, stride=32, bias=False)
        with torch.no_grad():
            filt = bilinear_filter(64)
            if self.grouped:
                self.upscore.weight.copy_(filt.expand(D, 1, 64, 64))
            else:
                self.upscore.weight.zero_()
                idx = torch.arange(D)
                self.upscore.weight[idx, idx] = filt
            self.seenmask_upscore.weight.zero_()
            self.seenmask_upscore.weight[torch.arange(2), torch.arange(2)] = filt
        self.upscore.weight.requires_grad = False  # never optimised by the reference (train.py:324-327)

    def forward(self, x):
        H, W = x.shape[2:]
        h = self.trunk(x)
        f = self.upscore(self.score_fr(h))[:, :, 19:19 + H, 19:19 + W].contiguous()
        s = self.seenmask_upscore(self.seenmask_score(h))[:, :, 19:19 + H, 19:19 + W].contiguous()  # both heads always run
        return f, s


def cosine_loss(score, target, table):
    """utils.py:75-102 for any batch size, target vectors gathered from the table."""
    te = table[target.clamp(min=0)].permute(0, 3, 1, 2)
    cos = (F.normalize(score.float(), dim=1) * F.normalize(te, dim=1)).sum(1)
    valid = target >= 0
    n = valid.sum()
    return (n - cos[valid].sum()) / n


def infer_lbl(score, table):
    n, c, h, w = score.shape
    en = table.norm(dim=1)
    en = torch.where(en == 0, torch.ones_like(en), en)
    s = score.float().permute(0, 2, 3, 1).reshape(-1, c)
    sim = (s @ table.t()) / (s.norm(dim=1, keepdim=True) * en[None])
    return sim.argmax(1).view(n, h, w)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=300)
    ap.add_argument("--C", type=int, default=59)
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--upscore", default="grouped", choices=["dense", "grouped"])
    ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"])
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(1337)
    m = StockFCN32s(a.D, a.upscore).to(dev).to(memory_format=torch.channels_last).train()
    x, lab, table = synth.synth_batch(a.B, 512, 512, a.C, a.D, seed=1337)
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    lab, table = lab.to(dev), table.to(dev)

    def step():
        m.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.dtype == "bf16"):
            f, s = m(x)
        loss = cosine_loss(f, lab, table)
        loss.backward()
        with torch.no_grad():
            lbl = infer_lbl(f.detach(), table)
        return loss, lbl

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss, _ = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"impl": "library (stock torch %s: cuDNN/cuBLAS)" % torch.__version__, "upscore": a.upscore,
                      "dtype": a.dtype, "B": a.B, "D": a.D, "C": a.C, "ms_per_step": ms,
                      "value": a.B * 512 * 512 / 1e6 / (ms / 1e3), "unit": "Mpixel/s", "loss": float(loss.item()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)


if __name__ == "__main__":
    main()
