"""Image-sharded data parallelism for the SZN hot path: one process per GPU, weights replicated,
NCCL all-reduce (sum) on parameter gradients only, plus a 2-element all-reduce of the loss accumulator
``[sum, n_valid]`` because the reference normalises by the number of valid pixels of the WHOLE batch
(``utils.py:95-101`` cosine, ``:46-47`` CE mean), so each rank must divide by the global count.

The reference has no distributed code; this is the one strategy the north star adds.
Gradients are reduced as they become final inside ``FCN32sFunction.backward`` (fc6's 411 MB first), on
NCCL's own stream, so the transfers overlap the remaining dgrad/wgrad kernels; tensors below
``small_bytes`` are packed into one flat buffer and reduced at the end.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class GradientAllReduce:
    """Attach to an ``FCN32s``: ``GradientAllReduce(model).accum_hook`` goes to the loss functions."""

    def __init__(self, model, group=None, small_bytes=1 << 20, sync_params=True):
        self.model = model
        self.group = group
        self.small_bytes = small_bytes
        self.works = []
        self.small = []
        self.bytes_reduced = 0
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        if self.enabled:
            model._grad_ready = self._on_ready
            model._grad_flush = self._flush
            if torch.cuda.is_available():
                # the all-reduce kernels share the SMs with the backward pass: finer split-K work items let the dynamic
                # tile scheduler route around busy SMs (include/szn.h: szn_set_wgrad_waves)
                from . import _lib
                _lib.call("szn_set_wgrad_waves", int(os.environ.get("SZN_DDP_WGRAD_WAVES", "3")))
            if sync_params:
                self.broadcast_parameters()

    def broadcast_parameters(self, src=0):
        """Replicas must start from the same weights: score_fr / seenmask_score keep torch's random default init
        (models.py:104-108), which differs per process unless every rank seeds identically.  Rank ``src``'s parameters and
        buffers are broadcast in place (channels_last parameters through their dense storage order)."""
        if not self.enabled:
            return
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                if t.is_contiguous():
                    dist.broadcast(t, src=src, group=self.group)
                elif t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous():
                    dist.broadcast(t.permute(0, 2, 3, 1), src=src, group=self.group)
                else:
                    tmp = t.contiguous()
                    dist.broadcast(tmp, src=src, group=self.group)
                    t.copy_(tmp)
        if hasattr(self.model, "_packed"):
            self.model._packed.invalidate()  # in-place writes under no_grad do bump versions; be explicit anyway

    def all_reduce_hist(self, hist):
        """Sum a device confusion histogram over the ranks (validation metrics must not be rank-local)."""
        if self.enabled:
            dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=self.group)
        return hist

    @property
    def is_main(self):
        """True on the one rank that may write logs / checkpoints."""
        return self.rank == 0

    def barrier(self):
        if self.enabled:
            dist.barrier(group=self.group)

    def _on_ready(self, name, g):
        nbytes = g.numel() * g.element_size()
        self.bytes_reduced += nbytes
        if nbytes < self.small_bytes:
            self.small.append(g)
            return
        if not g.is_contiguous():
            # conv weight gradients are OIHW-shaped views of a dense [O][R][S][I] buffer: reduce that buffer in place
            g = g.permute(0, 2, 3, 1)
            if not g.is_contiguous():
                self.small.append(g.permute(0, 3, 1, 2))
                return
        self.works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _flush(self):
        if self.small:
            flat = torch.cat([g.reshape(-1) for g in self.small])  # reshape copies the few strided ones
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            off = 0
            for g in self.small:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            self.small = []
        for w in self.works:
            w.wait()  # the compute stream waits on NCCL's stream; the host does not block
        self.works = []

    def accum_hook(self, accum):
        """In-place all-reduce of the loss accumulator [sum, n_valid] (fp64, device)."""
        if self.enabled:
            dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=self.group)

    def detach(self):
        self.model._grad_ready = None
        self.model._grad_flush = None


def shard_batch(n_items, rank, world):
    """Contiguous, balanced split of ``n_items`` images over ``world`` ranks -> (start, stop) of ``rank``."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
