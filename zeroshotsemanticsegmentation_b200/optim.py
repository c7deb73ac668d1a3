"""Fused optimizer steps: SGD for the FCN phase (``train.py:126-129``: ``torch.optim.SGD(params, lr, momentum=.99,
weight_decay=0.0005)`` with a second group for the biases at ``lr * 2, weight_decay 0``).

Same constructor, param-group and ``state_dict`` layout as ``torch.optim.SGD`` (the momentum buffer is stored under
``'momentum_buffer'``), so optimizer checkpoints of the reference (``trainer_fcn.py:281-292``) load unchanged; each
parameter is updated by ONE pass of ``szn_sgd_step`` over its storage (read p, g, buf; write p, buf) instead of torch's
three elementwise passes.  ``FusedAdam`` is ``torch.optim.Adam`` (``train.py:130-133`` optional FCN optimizer,
``train.py:175`` the seen-mask phase) the same way: ``szn_adam_step`` replaces six elementwise passes, state keys
``'step'`` / ``'exp_avg'`` / ``'exp_avg_sq'`` as torch stores them.  Conv weights live in channels_last memory and so do their gradients and buffers: the kernel
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
                if g.stride() != p.stride() or g.dtype != torch.float32 or g.data_ptr() % 16:
                    # same memory order as the parameter; 16-byte aligned (autograd may hand over a slice of a shared
                    # buffer, e.g. seenmask_score.bias.grad = dbh[D:D+2] with D % 4 != 0)
                    g = torch.empty_like(p).copy_(g)
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


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` with amsgrad off (all the reference uses)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("amsgrad is never enabled by the reference (train.py:133,175)")
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        st = None
        touched = []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam updates fp32 CUDA parameters only (no CPU fallback)")
                g = p.grad
                if g.stride() != p.stride() or g.dtype != torch.float32 or g.data_ptr() % 16:
                    # same memory order as the parameter; 16-byte aligned (autograd may hand over a slice of a shared
                    # buffer, e.g. seenmask_score.bias.grad = dbh[D:D+2] with D % 4 != 0)
                    g = torch.empty_like(p).copy_(g)
                state = self.state[p]
                if "exp_avg" not in state:
                    state["step"] = torch.tensor(0.0)  # host scalar tensor, as torch.optim.Adam keeps it
                    state["exp_avg"] = torch.zeros_like(p)  # preserve_format: the parameter's strides
                    state["exp_avg_sq"] = torch.zeros_like(p)
                for k in ("exp_avg", "exp_avg_sq"):
                    if state[k].stride() != p.stride():
                        state[k] = torch.empty_like(p).copy_(state[k])
                state["step"] += 1
                t = float(state["step"])
                step_size = group["lr"] / (1.0 - b1 ** t)
                bc2_sqrt = (1.0 - b2 ** t) ** 0.5
                if st is None:
                    st = _lib.stream()
                call("szn_adam_step", ptr(p), ptr(g), ptr(state["exp_avg"]), ptr(state["exp_avg_sq"]), p.numel(),
                     float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]), float(step_size),
                     float(bc2_sqrt), st)
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
        try:
            setter(tuple(params), tuple(p._version + 1 for p in params))
            return
        except TypeError:  # torch 2.x before the tuple overload: the signature is (Tensor, int)
            try:
                for p in params:
                    setter(p, p._version + 1)
                return
            except TypeError:
                pass
    for p in params:  # no usable setter: an in-place no-op does the same at the price of one more pass
        p.add_(0.0)
