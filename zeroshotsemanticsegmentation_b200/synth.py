"""Synthetic PASCAL-Context-shaped batches for bench.py: there is no network for datasets.

image  ~ U{0..255} minus the BGR mean (``pascal_dataset.py:39,138-145``), NCHW fp32;
labels ~ U{0..C-1} in ``block`` x ``block`` constant patches (segment-like), ``ignore_frac`` of the pixels set
         to -1 (the ignore label, ``pascal_dataset.py:120``), int64;
table  ~ N(0,1) (C, D), divided by its largest row norm (what the shipped ``norm_embed_arr_*.pkl`` are).
"""
import torch

MEAN_BGR = (104.00698793, 116.66876762, 122.67891434)


def synth_batch(B, H, W, C, D, seed=1337, block=32, ignore_frac=0.05):
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, 3, H, W), generator=g).float()
    x = img - torch.tensor(MEAN_BGR).view(1, 3, 1, 1)
    hb, wb = (H + block - 1) // block, (W + block - 1) // block
    lab = torch.randint(0, C, (B, hb, wb), generator=g)
    lab = lab.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()
    lab[torch.rand(B, H, W, generator=g) < ignore_frac] = -1
    table = torch.randn(C, D, generator=g)
    table = table / table.norm(dim=1).max()
    return x, lab, table


def init_model_(model, seed=1337):
    """Seeded default-torch conv init (VGG16 caffe weights need the network); deconvs keep the bilinear filter."""
    import math
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "upscore" in name:
                continue
            mod = dict(model.named_modules())[name.rsplit(".", 1)[0]]
            fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
            # He-style gain keeps activations O(1) through 15 ReLU layers, so the synthetic loss/gradients are not
            # denormal-small the way 1/sqrt(fan_in) uniform init makes them at depth
            bound = math.sqrt(6.0 / fan_in) if name.endswith("weight") else 1.0 / math.sqrt(fan_in)
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)
    return model
