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
    """Attach to an ``FCN32s``: ``GradientAllReduce(model).accum_hook`` goes to the loss functions.

    On CUDA the exchange runs through the C ABI (``szn_allreduce_bucket``, include/szn.h) on a communicator of its own:
    gradients are collected into buckets in the order the backward pass finishes them (heads + fc7 | fc6 | conv5 + conv4 |
    ... ) and each bucket is ONE fused NCCL launch on a side stream, so 32 parameter tensors cost a handful of launches
    and the transfers overlap the remaining dgrad / wgrad kernels.  Without CUDA (gloo, the CPU tests) every tensor goes
    through ``torch.distributed.all_reduce``."""

    def __init__(self, model, group=None, small_bytes=1 << 20, sync_params=True, bucket_bytes=32 << 20, use_c_abi=True):
        self.model = model
        self.group = group
        self.small_bytes = small_bytes
        self.bucket_bytes = bucket_bytes
        self.works = []
        self.small = []
        self.bucket, self.bucket_fill, self.fixups, self.launches = [], 0, [], 0
        self.bytes_reduced = 0
        self.comm = None
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.enabled else 0
        if self.enabled:
            model._grad_ready = self._on_ready
            model._grad_flush = self._flush
            if torch.cuda.is_available() and "nccl" in str(dist.get_backend(group)):
                from . import _lib
                # split-K work items per CTA of the weight gradient under data parallelism (include/szn.h)
                _lib.call("szn_set_wgrad_waves", int(os.environ.get("SZN_DDP_WGRAD_WAVES", "1")))
                if use_c_abi and _lib.load().szn_comm_available():
                    self._init_comm()
            if sync_params:
                self.broadcast_parameters()

    def _init_comm(self):
        """Rank 0 draws the NCCL unique id, torch.distributed ships it, every rank joins (szn_comm_init)."""
        import ctypes
        from . import _lib
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0:
            _lib.call("szn_comm_unique_id", buf)
        box = [buf.raw if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=self.group)
        comm = ctypes.c_void_p()
        _lib.call("szn_comm_init", box[0], self.rank, dist.get_world_size(self.group), ctypes.byref(comm))
        self.comm = comm
        self.comm_stream = torch.cuda.Stream()

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

    # ---- gradients ----
    @staticmethod
    def _dense(g):
        """The dense buffer behind a gradient: itself, or (conv weights) the [O][R][S][I] buffer its OIHW view permutes."""
        if g.is_contiguous():
            return g
        if g.dim() == 4 and g.permute(0, 2, 3, 1).is_contiguous():
            return g.permute(0, 2, 3, 1)
        return None

    def _on_ready(self, name, g):
        nbytes = g.numel() * g.element_size()
        self.bytes_reduced += nbytes
        if self.comm is not None:
            d = self._dense(g)
            if d is None or d.dtype != torch.float32:  # rare: reduce a dense fp32 copy, write it back when the backward ends
                d = g.float().contiguous()
                self.fixups.append((g, d))
            self.bucket.append(d)
            self.bucket_fill += nbytes
            if self.bucket_fill >= self.bucket_bytes:
                self._launch_bucket()
            return
        if nbytes < self.small_bytes:
            self.small.append(g)
            return
        d = self._dense(g)
        if d is None:
            self.small.append(g)
            return
        self.works.append(dist.all_reduce(d, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _launch_bucket(self):
        if not self.bucket:
            return
        import ctypes
        from . import _lib
        n = len(self.bucket)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in self.bucket])
        counts = (ctypes.c_longlong * n)(*[t.numel() for t in self.bucket])
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)  # the kernels that wrote these gradients have been enqueued on `cur`
        _lib.call("szn_allreduce_bucket", self.comm, ptrs, counts, n, _lib.F32, self.comm_stream.cuda_stream)
        for t in self.bucket:
            t.record_stream(self.comm_stream)
        self.launches += 1
        self.bucket, self.bucket_fill = [], 0

    def _flush(self):
        if self.comm is not None:
            self._launch_bucket()
            torch.cuda.current_stream().wait_stream(self.comm_stream)  # the host does not block
            for g, d in self.fixups:
                g.copy_(d.view_as(g))
            self.fixups = []
            return
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
        if not self.enabled:
            return
        if self.comm is not None and accum.is_cuda and accum.dtype == torch.float64 and accum.is_contiguous():
            import ctypes
            from . import _lib
            ptrs = (ctypes.c_void_p * 1)(accum.data_ptr())
            counts = (ctypes.c_longlong * 1)(accum.numel())
            _lib.call("szn_allreduce_bucket", self.comm, ptrs, counts, 1, 3, torch.cuda.current_stream().cuda_stream)
            return
        dist.all_reduce(accum, op=dist.ReduceOp.SUM, group=self.group)

    def detach(self):
        self.model._grad_ready = None
        self.model._grad_flush = None
        if self.comm is not None:
            from . import _lib
            torch.cuda.synchronize()
            _lib.call("szn_comm_destroy", self.comm)
            self.comm = None


def shard_batch(n_items, rank, world):
    """Contiguous, balanced split of ``n_items`` images over ``world`` ranks -> (start, stop) of ``rank``."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
