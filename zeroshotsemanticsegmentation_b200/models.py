"""Drop-in for the reference's ``models.py``: ``FCN32s`` with the same constructor, attribute names,
``state_dict`` keys/shapes and ``forward(x, mode)`` contract (``models.py:27-160``), executed by the
sm_100a kernels of ``libszn.so``.  There is no CPU / eager fallback.

Sub-modules are real ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.ReLU`` / ``nn.MaxPool2d`` /
``nn.Dropout2d`` instances so that ``train.py:get_parameters`` (``train.py:302-331``),
``copy_params_from_vgg16`` and checkpoint loading work unchanged; they only hold parameters —
``forward`` never calls them.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import engine


def bilinear_filter(kernel_size: int = 64) -> torch.Tensor:
    """2-D tent filter used for the upsampling deconvs (reference ``get_upsampling_weight``, models.py:11-19)."""
    factor = (kernel_size + 1) // 2
    center = factor - 1 if kernel_size % 2 == 1 else factor - 0.5
    k = np.arange(kernel_size, dtype=np.float64)
    w1 = 1.0 - np.abs(k - center) / factor
    return torch.from_numpy(w1[:, None] * w1[None, :]).float()


def get_upsampling_weight(in_channels, out_channels, kernel_size):
    """(in,out,k,k) weight with the bilinear filter on the channel diagonal (models.py:11-24)."""
    w = torch.zeros(in_channels, out_channels, kernel_size, kernel_size)
    n = min(in_channels, out_channels)
    idx = torch.arange(n)
    w[idx, idx] = bilinear_filter(kernel_size)
    return w


class FCN32s(nn.Module):
    """VGG16-FCN32s with an embedding head (``score_fr``/``upscore``) and a seen-mask head
    (``seenmask_score``/``seenmask_upscore``), reference ``models.py:27-160``.

    Extra keyword arguments (not in the reference):
      precision: ``"fp32"`` -- fp32-grade: every stored value is a (hi, lo) pair of bf16 planes (the 4 bytes of an
                 fp32) and every product is the error-compensated sum hi*hi + lo*hi + hi*lo of three bf16 tensor-core
                 MMAs with fp32 accumulate (~2^-16 per product): forward AND gradients match the reference's fp32 CPU
                 path to 1e-5 .. 1e-3 at every size, for 1.5x the tensor time of "tf32";
                 ``"tf32"`` (default) -- fp32 storage rounded to TF32, TF32 products (2^-11 per operand): the
                 throughput mode; forward within 1e-3 at the golden sizes, AT 1e-3 at 512x512 with O(1) activations;
                 ``"bf16"`` -- bf16 storage and products, fp32 accumulate.
      upscore_weight_grad: also compute the dense ``upscore.weight.grad`` (213 GFLOP/image at D=300 for a
                 weight the reference never optimises, ``train.py:324-327``); off by default.
      fused_head: the returned score carries a handle to the 17x17 score map it was upsampled from;
                 ``utils.cosine_loss / mse_loss(score, target, table=E)`` and ``utils.infer_lbl*`` then work from that
                 map (``szn_head_fused_*``: the x32 upsample is linear and per channel, so per-pixel dot products and
                 norms follow from a 289 x C product and a neighbour Gram matrix) instead of making four passes over the
                 (B, D, H, W) tensor, and the loss gradient re-enters the network as d s17.  Any other use of the score
                 (or an in-place change of it) silently takes the ordinary path.  Loss and gradients agree with the
                 ordinary path to fp32 rounding; labels differ only at near-ties (another summation order), so the
                 ordinary path stays the default and the parity reference.
    """

    def __init__(self, n_class=21, precision="tf32", upscore_weight_grad=False, fused_head=False):
        super().__init__()
        if precision not in engine.PRECISIONS:
            raise ValueError("precision must be one of %s" % list(engine.PRECISIONS))
        self.precision = precision
        self.upscore_weight_grad = upscore_weight_grad
        self.fused_head = bool(fused_head)
        for row in engine.TRUNK:
            if len(row) == 1:
                setattr(self, row[0], nn.MaxPool2d(2, stride=2, ceil_mode=True))
            else:
                name, cin, cout, k, pad = row
                setattr(self, name, nn.Conv2d(cin, cout, k, padding=pad))
                setattr(self, name.replace("conv", "relu"), nn.ReLU(inplace=True))
        self.fc6 = nn.Conv2d(512, 4096, 7)
        self.relu6 = nn.ReLU(inplace=True)
        self.drop6 = nn.Dropout2d()
        self.fc7 = nn.Conv2d(4096, 4096, 1)
        self.relu7 = nn.ReLU(inplace=True)
        self.drop7 = nn.Dropout2d()
        self.score_fr = nn.Conv2d(4096, n_class, 1)
        self.upscore = nn.ConvTranspose2d(n_class, n_class, 64, stride=32, bias=False)
        self.seenmask_score = nn.Conv2d(4096, 2, 1)
        self.seenmask_upscore = nn.ConvTranspose2d(2, 2, 64, stride=32, bias=False)
        self._initialize_weights()
        self._use_kernel_weight_layout()
        self._packed = engine.PackedWeights()
        self._grad_ready = None   # data-parallel hook: fn(name, grad) called as each gradient is final (ddp.py)
        self._grad_flush = None   # data-parallel hook: fn() called once at the end of backward
        self._forced_drop_masks = None  # tests inject Dropout2d masks here: (m6, m7), each (B,4096) in {0,1}

    def _initialize_weights(self):
        # convs keep torch's default init (the zero-init is commented out upstream, models.py:104-108)
        for m in self.modules():
            if isinstance(m, nn.ConvTranspose2d):
                assert m.kernel_size[0] == m.kernel_size[1]
                with torch.no_grad():
                    m.weight.copy_(get_upsampling_weight(m.in_channels, m.out_channels, m.kernel_size[0]))

    def _use_kernel_weight_layout(self):
        """Keep every k>1 conv weight in channels_last memory ([O][R][S][I] dense; shape, values and state_dict are
        unchanged).  That is the layout the wgrad kernel writes, so ``weight.grad`` is adopted by autograd as a view of
        the kernel's output buffer instead of being transposed through an extra pass over 135 M gradients."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d) and m.kernel_size[0] > 1 and m.in_channels > 3:
                m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)

    def invalidate_weight_cache(self):
        """Forget the packed (kernel-layout) copies of the weights.  Needed after editing parameters through ``.data``
        (``p.data.copy_()`` / ``mul_()`` / ``clamp_()``), which does not advance the version counters the cache keys on."""
        self._packed.invalidate()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_packed"):
            self._packed.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._packed.invalidate()
        return out

    def _ordered_params(self):
        sd = dict(self.named_parameters())
        return [sd[n] for n in engine.PARAM_ORDER]

    def forward(self, x, mode="fcn"):
        if mode not in ("fcn", "seenmask", "both"):
            raise Exception("model given unexpected forward mode")
        if not x.is_cuda:
            raise RuntimeError("FCN32s (B200 build) runs on CUDA only: move the model and input to the GPU")
        self._grad_enabled = torch.is_grad_enabled()  # the autograd function's forward always runs with grad mode off
        out = engine.FCN32sFunction.apply(self, x, *self._ordered_params())
        f, s = out[0], out[1]
        if len(out) == 3:  # fused_head: the score remembers the map it came from (see utils.ScoreHandle)
            f._szn_head = engine.ScoreHandle(out[2], f)
        if mode == "fcn":
            return f
        if mode == "seenmask":
            return s
        return f, s

    def copy_params_from_vgg16(self, vgg16):
        """models.py:162-193: conv weights from ``vgg16.features``, fc6/fc7 from ``classifier[0|3]``."""
        mine = [getattr(self, r[0]) for r in engine.TRUNK if len(r) == 5]
        theirs = [m for m in vgg16.features if isinstance(m, nn.Conv2d)]
        for src, dst in zip(theirs, mine):
            assert src.weight.size() == dst.weight.size() and src.bias.size() == dst.bias.size()
            dst.weight.data = src.weight.data
            dst.bias.data = src.bias.data
        for i, name in zip([0, 3], ["fc6", "fc7"]):
            src, dst = vgg16.classifier[i], getattr(self, name)
            dst.weight.data = src.weight.data.view(dst.weight.size())
            dst.bias.data = src.bias.data.view(dst.bias.size())
        self._use_kernel_weight_layout()  # the copied tensors are NCHW-dense: back to the layout the wgrad kernel writes
        self._packed.invalidate()         # `.data =` swaps storage without touching the version counters
