"""world_size-2 gloo tests of the multi-GPU host logic (batch sharding + logits gather) on CPU.
The compute is stubbed by a deterministic per-image function: what is under test is the plumbing of
lang-seg_b200/parallel.py (rank order, ragged shards, shapes), which is backend-independent."""
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


class _FakeNet:
    """Stands in for LSegNet.forward: per-image logits that depend only on that image."""

    def __call__(self, x, labelset=""):
        k = 3
        s = x.flatten(1).sum(dim=1).view(-1, 1, 1, 1)
        base = torch.arange(k, dtype=torch.float32).view(1, k, 1, 1)
        return (s + base).expand(x.shape[0], k, 4, 4).contiguous()


def _worker(rank, world, port, batch, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lseg_b200  # noqa: F401
        from lseg_b200 import parallel
        g = torch.Generator().manual_seed(0)
        x = torch.randn(batch, 3, 8, 8, generator=g)
        out = parallel.forward_sharded(_FakeNet(), x)
        ref = _FakeNet()(x)
        ok = out.shape == ref.shape and torch.equal(out, ref)
        lo, hi = parallel.shard_bounds(batch, rank, world)
        ret[rank] = (bool(ok), lo, hi)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5])
def test_sharded_forward_gloo_world2(batch):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]
    assert ret[0][1] == 0 and ret[0][2] == ret[1][1] and ret[1][2] == batch


def test_shard_bounds_cover_batch():
    import lseg_b200  # noqa: F401
    from lseg_b200.parallel import shard_bounds
    for batch in (1, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
