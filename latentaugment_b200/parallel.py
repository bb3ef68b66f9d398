"""Multi-GPU plumbing (one process per GPU, ``torch.distributed``).

The augmentation loop shards by batch with NO data-path collective (samples are independent
given G).  The only exchange on the path is the nearest-code query against a
bank sharded by rows: every rank holds all queries, searches its shard (tap-GEMM + fused top-k +
exact re-rank, csrc/distance.cu) and the per-shard ``(distance, global index)`` lists are merged
after one ``all_gather`` of ``n*k*12`` bytes per rank (NCCL over NVLink on GPUs).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous balanced shard [begin, end) of n items for `rank` of `world`."""
    return n * rank // world, n * (rank + 1) // world


def all_gather_queries(x_local, group=None):
    """Gathers per-rank query shards [n_r, K] (equal n_r) into [world*n_r, K] on every rank."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(x_local) for _ in range(world)]
    dist.all_gather(out, x_local.contiguous(), group=group)
    return torch.cat(out)


def sharded_nearest_codes(nearest_fn, X, k, group=None, merge_fn=None):
    """``nearest_fn(X, k) -> (dist [n,k] f32, idx [n,k] i64 GLOBAL indices)`` on this rank's bank shard
    (``LatentBank(shard, index_offset=begin).nearest``).  Returns the global k best on every rank."""
    if merge_fn is None:
        from .engine import merge_topk as merge_fn
    d, i = nearest_fn(X, k)
    world = dist.get_world_size(group)
    ds = [torch.empty_like(d) for _ in range(world)]
    is_ = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(ds, d.contiguous(), group=group)
    dist.all_gather(is_, i.contiguous(), group=group)
    return merge_fn(torch.stack(ds), torch.stack(is_))
