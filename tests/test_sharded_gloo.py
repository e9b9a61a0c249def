"""CPU, world_size 2, gloo: host logic of the sharded layout (range split, all-gather, merge rule).
The local searcher here is the oracle over each rank's shard graph (test infrastructure); on the
GPU the same split runs as ShardedEngine: walk, one NCCL all-gather, merge kernel (tests/test_gpu_multi.py, 2 ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, Fixture


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle as O
    from annlite_b200.engine import Engine
    from annlite_b200.sharded import ShardedSearcher, merge_topk_host, shard_range
    fx = Fixture('l2_m4')
    n = len(fx.codes)
    lo, hi = shard_range(n, rank, world)
    # each rank builds ITS shard graph with the product's host builder (host-only handle) ...
    X = fx.X[lo:hi]
    T = O.adc_table(X, fx.cb, fx.metric)
    e = Engine(fx.M * fx.ds, fx.M, fx.Ks, fx.metric, device=-1)
    e.init_graph(hi - lo, M=16, ef_construction=100)
    e.add_items_with_tables(fx.codes[lo:hi], T, fx.labels[lo:hi], num_threads=1)
    g = O.Graph.from_state(e.get_graph(), fx.M, fx.Ks)
    tq = O.adc_table(fx.Q, fx.cb, fx.metric)

    def local(queries, k):
        l, d, _ = O.hnsw_search(g, tq, k, fx.ef)
        return l.view(np.int64), d

    s = ShardedSearcher(local)
    ml, md = s.search(fx.Q, fx.k)
    # every rank ends with the same merged answer
    gl = [None] * world
    dist.all_gather_object(gl, (ml.tolist(), md.tolist()))
    assert gl[0] == gl[1]
    # and it equals merging the two shard results directly with the reference's rule
    ll, dd = local(fx.Q, fx.k)
    both = [None] * world
    dist.all_gather_object(both, (ll.view(np.uint64), dd))
    L = np.stack([b[0] for b in both])
    D = np.stack([b[1] for b in both])
    el, ed = merge_topk_host(L, D, fx.k)
    assert np.array_equal(ml, el) and np.array_equal(md, ed)
    # labels are global ids from both halves, distances ascending
    assert (np.diff(md, axis=1) >= 0).all()
    own = set(fx.labels[lo:hi].tolist())
    frac_own = np.mean([[int(x) in own for x in row] for row in ml])
    assert 0.1 < frac_own < 0.9
    if rank == 0:
        np.save(os.path.join(tmp, 'ok.npy'), np.array([1]))
    dist.destroy_process_group()


def test_sharded_search_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / 'ok.npy')


def test_ranges_cover_everything():
    from annlite_b200.sharded import query_slice, shard_range
    for n in (0, 1, 7, 1000, 10 ** 6 + 3):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            assert [query_slice(n, r, w) for r in range(w)] == edges


def test_merge_rule_matches_container_semantics():
    from annlite_b200.sharded import merge_topk_host
    rng = np.random.default_rng(0)
    G, B, k = 3, 5, 4
    d = np.sort(rng.random((G, B, k)).astype(np.float32), axis=2)
    l = rng.permutation(G * B * k).astype(np.uint64).reshape(G, B, k)
    d[1, 0, :] = d[0, 0, :]          # exact ties across shards -> ordered by label
    l[2, 1, 2:] = np.iinfo(np.uint64).max   # a shard that found fewer than k
    d[2, 1, 2:] = np.inf
    ml, md = merge_topk_host(l, d, k)
    for b in range(B):
        # container.py:130-138: hstack -> argsort[:limit]
        dd, ll = np.hstack(d[:, b]), np.hstack(l[:, b])
        keep = ll != np.iinfo(np.uint64).max
        dd, ll = dd[keep], ll[keep]
        order = np.lexsort((ll, dd))[:k]
        assert np.array_equal(md[b], dd[order]) and np.array_equal(ml[b], ll[order])
