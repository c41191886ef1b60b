"""world_size-2 gloo test of the N>1 host logic: contiguous query sharding + one
all-gather of per-rank top-k.  The per-rank "search" is the oracle here (no GPU in
this container); on the GPU box bench.py runs the same plumbing over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import queries_from, small_index


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ix, q, k, w, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from freddy_b200.dist import allgather_topk, shard_range
    from oracle import oracle
    b, e = shard_range(len(q), rank, world)
    ids, d, rc, _ = oracle.OracleIndex(ix).ivfadc_search(q[b:e], k, w)
    assert rc == 0
    gi, gd = allgather_topk(torch.from_numpy(ids), torch.from_numpy(d), len(q))
    if rank == 1:   # a non-zero rank reports, so the gather is checked where it matters
        out.put((gi.numpy(), gd.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from freddy_b200.dist import shard_range
    for n in (0, 1, 7, 10, 1001):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_matches_single(oracle_mod):
    ix = dict(small_index())
    ix.pop("vectors", None)
    q = queries_from(small_index(), 51)        # ragged: 26 + 25
    k, w = 5, 4
    eids, ed, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, k, w)
    assert rc == 0
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ix, q, k, w, out)) for r in range(2)]
    for p in procs:
        p.start()
    gi, gd = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(gi, eids)
    np.testing.assert_array_equal(gd.view(np.uint32), ed.view(np.uint32))


def _worker_vocab(rank, world, port, vectors, vec_ids, q, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from freddy_b200.dist import allgather_merge_topk_desc, shard_range
    from oracle import oracle
    b, e = shard_range(len(vectors), rank, world)                  # this rank's rows of the word-vector table
    ids, s = oracle.knn_exact(vectors[b:e], vec_ids[b:e], q, k)
    gi, gs = allgather_merge_topk_desc(torch.from_numpy(ids), torch.from_numpy(s))
    if rank == 1:
        out.put((gi.numpy(), gs.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_vocabulary_sharded_exact_knn(oracle_mod):
    """exact k-NN with the vocabulary split over two ranks + one all-gather == the single-rank scan, incl. duplicated
    vectors whose equal similarities straddle the shard boundary (global row order must survive the merge)"""
    ix = small_index()
    vectors = ix["vectors"][:3001].copy()
    vec_ids = np.asarray(ix["ids"][:3001], np.int32)
    vectors[2000:2010] = vectors[100:110]                          # duplicates in the other shard
    q = np.ascontiguousarray(vectors[[100, 105, 2500, 7]], np.float32)
    k = 6
    eids, es = oracle_mod.knn_exact(vectors, vec_ids, q, k)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_vocab, args=(r, 2, port, vectors, vec_ids, q, k, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    gi, gs = out.get(timeout=120)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    np.testing.assert_array_equal(gi, eids)
    np.testing.assert_array_equal(gs.view(np.uint32), es.view(np.uint32))
