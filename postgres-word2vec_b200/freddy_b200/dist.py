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


def allgather_merge_topk_desc(ids, sims, group=None):
    """Vocabulary-sharded exact k-NN (k_nearest_neighbour over a word-vector table split row-wise across ranks in
    rank order): ids/sims are this rank's [nq, k] best rows of ITS shard, ordered (similarity desc, row asc), id -1 =
    unfilled.  One all-gather, then every rank merges the world*k candidates of each query by
    (similarity desc, rank asc, local position asc) — which is (similarity desc, global row asc) because the shards
    are contiguous row ranges in rank order — and keeps the first k.  Returns ([nq, k] ids, [nq, k] sims)."""
    world = dist.get_world_size(group)
    nq, k = ids.shape
    all_i = torch.empty((world, nq, k), dtype=ids.dtype, device=ids.device)
    all_s = torch.empty((world, nq, k), dtype=sims.dtype, device=sims.device)
    dist.all_gather_into_tensor(all_i.view(world * nq, k), ids.contiguous(), group=group)
    dist.all_gather_into_tensor(all_s.view(world * nq, k), sims.contiguous(), group=group)
    cand_i = all_i.permute(1, 0, 2).reshape(nq, world * k)          # per query: rank-major, local order inside
    cand_s = all_s.permute(1, 0, 2).reshape(nq, world * k)
    key = torch.where(cand_i >= 0, cand_s.double(), torch.full_like(cand_s, float("-inf")).double())
    order = torch.sort(key, dim=1, descending=True, stable=True).indices[:, :k]     # stable: ties keep rank/local order
    out_i = torch.gather(cand_i, 1, order)
    out_s = torch.gather(cand_s, 1, order)
    return out_i, torch.where(out_i >= 0, out_s, torch.zeros_like(out_s))
