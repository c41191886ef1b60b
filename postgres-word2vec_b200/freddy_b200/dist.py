"""Multi-GPU plumbing for the query-sharded batch paths.

Queries are independent, so the index is replicated on every GPU and the query
array is sharded contiguously across ranks; the only exchange of the path is one
all-gather of the per-rank top-k `(id, distance)` tables.  One process per GPU,
torch.distributed (NCCL on GPUs; gloo works for the CPU tests of this logic).
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous, balanced [begin, end) of n items for `rank` of `world`"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allgather_topk(ids, dists, n_total, group=None):
    """ids/dists: this rank's [n_local, k] tensors (any device the backend supports).
    Returns the [n_total, k] tables in query order on every rank.  Shards may be
    ragged (n_total not divisible by world): pad to the largest shard, gather, trim."""
    world = dist.get_world_size(group)
    k = ids.shape[1]
    n_max = (n_total + world - 1) // world
    pad_i = torch.full((n_max, k), -1, dtype=ids.dtype, device=ids.device)
    pad_d = torch.zeros((n_max, k), dtype=dists.dtype, device=dists.device)
    pad_i[: ids.shape[0]] = ids
    pad_d[: dists.shape[0]] = dists
    out_i = torch.empty((world * n_max, k), dtype=ids.dtype, device=ids.device)
    out_d = torch.empty((world * n_max, k), dtype=dists.dtype, device=dists.device)
    dist.all_gather_into_tensor(out_i, pad_i, group=group)
    dist.all_gather_into_tensor(out_d, pad_d, group=group)
    parts_i, parts_d = [], []
    for r in range(world):
        b, e = shard_range(n_total, r, world)
        parts_i.append(out_i[r * n_max: r * n_max + (e - b)])
        parts_d.append(out_d[r * n_max: r * n_max + (e - b)])
    return torch.cat(parts_i), torch.cat(parts_d)
