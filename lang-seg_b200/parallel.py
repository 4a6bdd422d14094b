"""Multi-GPU host logic: one process per GPU, batch sharded, one collective.

The image path has no cross-image operation in eval mode (BatchNorm uses running statistics), so the
batch shards across ranks with NO data-path collective; weights are resident per rank and text features
are computed redundantly per rank (150 KB, cheaper than a broadcast). The only exchange is the gather of
the final logits: an NCCL all-gather over NVLink 5 / NVSwitch on B200 (gloo in the CPU tests).
This replaces the reference's single-process, thread-per-GPU DataParallel with a per-call parameter
broadcast (additional_utils/models.py:35-53,183-248).
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous shard [lo, hi) of a batch of `batch` images for `rank` (remainder to the low ranks)."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(x, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_logits(local, batch=None, group=None):
    """All-gather per-rank logits [b_r, K, H, W] into [sum b_r, K, H, W] in rank order on every rank.
    Equal shards use a single all_gather_into_tensor (in-place receive); ragged shards pad to the max."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    mx = max(sizes)
    padded = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(bufs, padded, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def forward_sharded(net, x_global, labelset=""):
    """Each rank runs net.forward on its contiguous shard of the global batch and gathers the logits."""
    local = shard_batch(x_global)
    out = net(local.contiguous(), labelset) if labelset != "" else net(local.contiguous())
    return gather_logits(out)
