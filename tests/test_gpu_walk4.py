"""GPU: hnsw_walk4 (walk_fused.cu) -- the plain search with K1 fused into the walk.

* the table the walk builds in shared memory is bit-equal to pq_bind's (oracle restatement of
  bindings/pq_bindings.pyx:149-274 + pq.py:316-322), for L2 / IP / cosine, V=4 and V=2 codebook layouts;
* the three kernels of the plain search -- hnsw_walk4 fused (default), hnsw_walk4 over materialised
  tables (TMA staging) and the round-1 hnsw_walk_fast -- return identical labels, distance bits and
  hop / neighbour counters on every query (they implement the same single-list walk);
* against the oracle (searchKnn restatement): identical on every walk that met no exact fp32 tie.
"""
import numpy as np
import pytest

import oracle as O
from annlite_b200.engine import Engine
from helpers import bits, tie_aware_rows

pytestmark = pytest.mark.gpu


def make(N, D, M, metric, seed, Mconn=16, efc=100, nq=200):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    ds = D // M
    Xi = O.l2_normalize(X).astype(np.float32) if metric == 'cosine' else X
    cb = np.stack([Xi[rng.choice(N, 256, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, 256, metric)
    e.set_codebook(cb)
    e.init_graph(N, M=Mconn, ef_construction=efc)
    e.add_items(Xi, np.arange(N, dtype=np.uint64) + 5)
    g = O.Graph.from_state(e.get_graph(), M, 256)
    return e, g, cb, Q


def oracle_tables(Q, cb, metric):
    if metric == 'cosine':   # HnswIndex.search: pre_process normalises, get_dist_mat normalises again
        Q = O.l2_normalize(Q).astype(np.float32)
    return O.adc_table(Q, cb, metric)


@pytest.mark.parametrize('D,M,metric', [(128, 8, 'euclidean'), (96, 16, 'euclidean'), (768, 32, 'euclidean'),
                                        (128, 8, 'inner_product'), (48, 8, 'euclidean'), (64, 32, 'inner_product')])
def test_fused_table_is_bit_equal_to_pq_bind(D, M, metric):
    """dump_tables exports what the walk had in shared memory (ds=16,6,24: V=4,2,4; ds=6 and 2: V=2)."""
    import torch
    e, g, cb, Q = make(3000, D, M, metric, 21, nq=130)
    dump = torch.zeros((Q.shape[0], M, 256), dtype=torch.float32, device='cuda')
    e.set_option('dump_tables', dump.data_ptr())
    e.search(queries=Q, k=10, ef=64)
    e.set_option('dump_tables', 0)
    e.sync()
    t = O.adc_table(Q, cb, metric)
    assert np.array_equal(bits(dump.cpu().numpy()), bits(t))


def test_fused_table_cosine_within_tolerance():
    """cosine: the device normalises (einsum order differs from numpy's): <= 1e-5 rel, ids equal."""
    import torch
    e, g, cb, Q = make(3000, 128, 8, 'cosine', 22, nq=100)
    dump = torch.zeros((Q.shape[0], 8, 256), dtype=torch.float32, device='cuda')
    e.set_option('dump_tables', dump.data_ptr())
    l, d = e.search(queries=Q, k=10, ef=64, normalize=2)
    e.set_option('dump_tables', 0)
    t = oracle_tables(Q, cb, 'cosine')
    got = dump.cpu().numpy()
    assert np.allclose(got, t, rtol=1e-5, atol=1e-6)
    ol, od, found = O.hnsw_search(g, t, 10, 64)
    assert (l == ol).mean() > 0.995


@pytest.mark.parametrize('D,M,k,ef,Mconn', [(128, 8, 10, 64, 16), (128, 8, 1, 1, 16), (128, 8, 10, 10, 16),
                                            (128, 8, 33, 50, 16), (128, 8, 100, 100, 16), (128, 8, 10, 200, 16),
                                            (128, 8, 10, 500, 16), (96, 16, 10, 128, 16), (96, 16, 10, 256, 16),
                                            (768, 32, 100, 128, 16), (128, 8, 10, 64, 8), (128, 8, 10, 64, 5)])
def test_three_kernels_agree_and_match_the_oracle(D, M, k, ef, Mconn):
    N = 20000 if D <= 128 else 6000
    e, g, cb, Q = make(N, D, M, 'euclidean', 100 + ef + k, Mconn=Mconn)
    t = O.adc_table(Q, cb, 'euclidean')
    res = {}
    for wk in (0, 2, 1):
        e.set_option('walk_kernel', wk)
        res[wk] = e.search(queries=Q, k=k, ef=ef, with_stats=True)
    e.set_option('walk_kernel', 0)
    res['tables'] = e.search(tables=t, k=k, ef=ef, with_stats=True)     # literal dtables form -> TMA staging
    for other in (2, 1, 'tables'):
        assert np.array_equal(res[0][0], res[other][0]), other
        assert np.array_equal(bits(res[0][1]), bits(res[other][1])), other
        assert np.array_equal(res[0][2], res[other][2]), other            # hops, neighbours, evaluations
    l, d, st = res[0]
    ol, od, found, (hops, nbrs, evals), ties = O.hnsw_search(g, t, k, ef, with_counts=True, with_ties=True)
    assert (found == min(k, N)).all()
    v = np.array(tie_aware_rows(l, d, ol, od))
    clean = ties == 0
    assert (v[clean] == 'exact').all(), (v[clean] != 'exact').sum()
    assert np.array_equal(st[clean, 0], hops[clean]) and np.array_equal(st[clean, 1], nbrs[clean])
    assert (v[~clean] == 'diff').sum() <= 1
    # and bit for bit the scalar model of the single-list walk, tie rows included
    ml, md, mfound, mhops, mnbrs = O.single_list_walk(g, t, k, ef)
    assert np.array_equal(l, ml) and np.array_equal(bits(d), bits(md))
    assert np.array_equal(st[:, 0], mhops) and np.array_equal(st[:, 1], mnbrs)


def test_tiny_graphs_and_batches():
    for N in (1, 2, 3, 17, 40):
        rng = np.random.default_rng(300 + N)
        X = rng.standard_normal((N, 128)).astype(np.float32)
        Q = rng.standard_normal((7, 128)).astype(np.float32)
        cb = rng.standard_normal((8, 256, 16)).astype(np.float32)
        e = Engine(128, 8, 256, 'euclidean')
        e.set_codebook(cb)
        e.init_graph(N, M=16, ef_construction=50)
        e.add_items(X, np.arange(N, dtype=np.uint64))
        g = O.Graph.from_state(e.get_graph(), 8, 256)
        t = O.adc_table(Q, cb, 'euclidean')
        k = min(N, 3)
        l, d, st = e.search(queries=Q, k=k, ef=8, with_stats=True)
        ol, od, found, (hops, nbrs, evals) = O.hnsw_search(g, t, k, 8, with_counts=True)
        assert np.array_equal(l, ol) and np.array_equal(bits(d), bits(od))
        assert np.array_equal(st[:, 0], hops)
        for B in (1, 2, 5):     # fewer queries than SMs
            l2, d2 = e.search(queries=Q[:B], k=k, ef=8)
            assert np.array_equal(l2, ol[:B])


def test_streamed_and_chunked_forms_use_the_fused_kernel():
    import torch
    e, g, cb, Q = make(20000, 128, 8, 'euclidean', 77, nq=9000)
    t = O.adc_table(Q[:500], cb, 'euclidean')
    ol, od, found = O.hnsw_search(g, t, 10, 64)
    l, d = e.search(queries=Q, k=10, ef=64)                       # >= 4096 host queries: chunked two-stream pipeline
    assert np.array_equal(l[:500], ol) and np.array_equal(bits(d[:500]), bits(od))
    Qd = torch.from_numpy(Q).cuda()
    outs = [(torch.empty((9000, 10), dtype=torch.int64, device='cuda'), torch.empty((9000, 10), dtype=torch.float32, device='cuda'))
            for _ in range(2)]
    n0 = e.launch_count
    tk = [e.search_submit(Qd, outs[i][0], outs[i][1], k=10, ef=64) for i in range(2)]
    for x in tk:
        e.search_wait(x)
    assert e.launch_count - n0 == 2                                # one kernel per batch: no K1 launch
    for i in range(2):
        assert np.array_equal(outs[i][0].cpu().numpy().view(np.uint64), l)
        assert np.array_equal(bits(outs[i][1].cpu().numpy()), bits(d))
