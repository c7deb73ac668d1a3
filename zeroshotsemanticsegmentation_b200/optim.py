"""Fused SGD step for the FCN phase (``train.py:126-129``: ``torch.optim.SGD(params, lr, momentum=.99,
weight_decay=0.0005)`` with a second group for the biases at ``lr * 2, weight_decay 0``).

Same constructor, param-group and ``state_dict`` layout as ``torch.optim.SGD`` (the momentum buffer is stored under
``'momentum_buffer'``), so optimizer checkpoints of the reference (``trainer_fcn.py:281-292``) load unchanged; each
parameter is updated by ONE pass of ``szn_sgd_step`` over its storage (read p, g, buf; write p, buf) instead of torch's
three elementwise passes.  Conv weights live in channels_last memory and so do their gradients and buffers: the kernel
walks raw storage, which is valid because all three share strides.  SURVEY §8f row 3.
"""
import torch

from . import _lib
from ._lib import call, ptr


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        if lr < 0 or momentum < 0 or weight_decay < 0:
            raise ValueError("lr, momentum and weight_decay must be non-negative")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = None
        touched = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedSGD updates fp32 CUDA parameters only (no CPU fallback)")
                g = p.grad
                if g.stride() != p.stride() or g.dtype != torch.float32:
                    g = torch.empty_like(p).copy_(g)  # same memory order as the parameter
                state = self.state[p]
                first = "momentum_buffer" not in state or state["momentum_buffer"] is None
                if first:
                    state["momentum_buffer"] = torch.empty_like(p)  # preserve_format: the parameter's strides
                buf = state["momentum_buffer"]
                if buf.stride() != p.stride():
                    buf = state["momentum_buffer"] = torch.empty_like(p).copy_(buf)
                if st is None:
                    st = _lib.stream()
                call("szn_sgd_step", ptr(p), ptr(g), ptr(buf), p.numel(), float(group["lr"]), float(group["momentum"]),
                     float(group["weight_decay"]), int(first), st)
                touched.append(p)
        _bump_versions(touched)
        return loss


def _bump_versions(params):
    """The kernel wrote the parameters behind autograd's back: advance their version counters, which is what the
    module's packed-weight cache (engine.PackedWeights) and autograd's saved-tensor checks key on."""
    if not params:
        return
    setter = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
    if setter is not None:
        setter(tuple(params), tuple(p._version + 1 for p in params))
    else:  # older torch: an in-place no-op does the same at the price of one more pass
        for p in params:
            p.add_(0.0)
