"""Host-side logic of the image-sharded data parallelism (ddp.py) on CPU: world_size 2, gloo backend.
The gradient buckets, the strided (channels_last) weight-gradient views and the loss-accumulator hook must give every
rank the SUM over ranks, which is what makes N-GPU training equal to single-GPU training on the concatenated batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _FakeModel:
    _grad_ready = None
    _grad_flush = None

    def parameters(self):
        return []

    def buffers(self):
        return []


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zeroshotsemanticsegmentation_b200 import ddp
    torch.manual_seed(100 + rank)
    m = _FakeModel()
    red = ddp.GradientAllReduce(m, small_bytes=1 << 12)
    assert red.enabled and m._grad_ready is not None
    big = torch.randn(64, 3, 3, 32)                      # dense [O][R][S][I] buffer written by the wgrad kernel
    g_big = big.permute(0, 3, 1, 2)                      # OIHW-shaped strided view handed to autograd
    g_small = torch.randn(64)                            # a bias gradient (goes to the flat bucket)
    g_small_strided = torch.randn(8, 2, 2, 4).permute(0, 3, 1, 2)  # small AND strided
    g_plain = torch.randn(300, 40)                       # contiguous, above the bucket threshold
    local = [t.clone() for t in (g_big, g_small, g_small_strided, g_plain)]
    for name, g in (("conv.weight", g_big), ("conv.bias", g_small), ("c2.weight", g_small_strided), ("fc.weight", g_plain)):
        m._grad_ready(name, g)
    m._grad_flush()
    accum = torch.tensor([1.5 + rank, 10.0 * (rank + 1)], dtype=torch.float64)
    red.accum_hook(accum)
    gathered = [None] * world
    dist.all_gather_object(gathered, [t.contiguous() for t in local])
    want = [sum(g[i] for g in gathered) for i in range(4)]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip((g_big, g_small, g_small_strided, g_plain), want))
    ok = ok and torch.allclose(accum, torch.tensor([1.5 + 2.5, 10.0 + 20.0], dtype=torch.float64))
    ok = ok and red.bytes_reduced == sum(t.numel() * 4 for t in local)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def _bcast_worker(rank, world, port, out):
    """Replicas start from rank 0's weights (score_fr / seenmask_score keep a per-process random init), also
    for channels_last parameters; validation histograms are summed over the ranks; only rank 0 is the writer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from zeroshotsemanticsegmentation_b200 import ddp
    torch.manual_seed(100 + rank)  # every rank draws other initial weights

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = torch.nn.Conv2d(8, 4, 3)
            self.conv.weight.data = self.conv.weight.data.contiguous(memory_format=torch.channels_last)
            self.fc = torch.nn.Linear(5, 3)
            self.register_buffer("filt", torch.randn(4, 4))
            self._grad_ready = self._grad_flush = None

    m = M()
    before = [p.detach().clone() for p in m.parameters()]
    red = ddp.GradientAllReduce(m)
    state = [t.detach().contiguous().clone() for t in list(m.parameters()) + list(m.buffers())]
    gathered = [None] * world
    dist.all_gather_object(gathered, state)
    same = all(torch.equal(a, b) for a, b in zip(gathered[0], gathered[1]))
    kept_layout = m.conv.weight.permute(0, 2, 3, 1).is_contiguous()
    rank0_untouched = rank != 0 or all(torch.equal(a, b) for a, b in zip(before, m.parameters()))
    hist = torch.full((3, 3), float(rank + 1))
    red.all_reduce_hist(hist)
    out[rank] = bool(same and kept_layout and rank0_untouched and torch.equal(hist, torch.full((3, 3), 3.0))
                     and red.is_main == (rank == 0))
    dist.destroy_process_group()


def test_parameters_are_broadcast_and_hists_reduced_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_bcast_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_gradient_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_shard_batch_partitions_images():
    from zeroshotsemanticsegmentation_b200.ddp import shard_batch
    for n, w in ((64, 8), (10, 4), (3, 8), (128, 8)):
        spans = [shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_single_process_is_a_no_op():
    from zeroshotsemanticsegmentation_b200 import ddp
    m = _FakeModel()
    red = ddp.GradientAllReduce(m)
    assert not red.enabled and m._grad_ready is None
    a = torch.tensor([1.0, 2.0], dtype=torch.float64)
    red.accum_hook(a)
    assert a.tolist() == [1.0, 2.0]
