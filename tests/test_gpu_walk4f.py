"""GPU: hnsw_walk4f (walk_flagged4.cu) -- the filtered / deletion-aware search on two register lists
(searchBaseLayerSTWithFilter hnswalg.h:332-440, searchBaseLayerST<true> :243-329, knn_query_with_filter
bindings/hnsw_bindings.cpp:393-516), with K1 fused into the walk.

* against the oracle (searchKnnWithFilter / deletion-aware searchKnn restatement): identical labels, distance
  bits, hop and neighbour counters on every walk that met no exact fp32 tie, for random filters of several
  selectivities, deletions, and the forced (no filter, no deletions) route;
* the fused form (queries in), the literal `tables=` form (TMA staging), round 1's hnsw_walk_flagged and the
  bitmap walk return the same rows;
* a traversed-only list that is too small on purpose: exactly the queries the scalar model flags are re-run on
  the bitmap walk, and every row is still the oracle's;
* the streamed form (annb_search_submit_filtered) returns what the blocking call returns.
"""
import numpy as np
import pytest

import oracle as O
from annlite_b200.engine import Engine
from helpers import bits, tie_aware_rows

pytestmark = pytest.mark.gpu


def make(N, D, M, seed, Mconn=16, efc=100, nq=200, metric='euclidean'):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    ds = D // M
    cb = np.stack([X[rng.choice(N, 256, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, 256, metric)
    e.set_codebook(cb)
    e.init_graph(N, M=Mconn, ef_construction=efc)
    labels = np.arange(N, dtype=np.uint64) * 3 + 5
    e.add_items(X, labels)
    return e, cb, Q, labels, rng


def check_vs_oracle(l, d, st, g, t, k, ef, allow):
    ol, od, found, (hops, nbrs, evals), ties = O.hnsw_search(g, t, k, ef, filter_labels=allow, with_counts=True, with_ties=True)
    assert (found == k).all()
    v = np.array(tie_aware_rows(l, d, ol, od))
    clean = ties == 0
    assert (v[clean] == 'exact').all(), int((v[clean] != 'exact').sum())
    assert (v[~clean] == 'diff').sum() <= 1
    if st is not None:
        assert np.array_equal(st[clean, 0], hops[clean]) and np.array_equal(st[clean, 1], nbrs[clean])
    return v


@pytest.mark.parametrize('D,M,k,ef,s', [(128, 8, 10, 64, 0.5), (128, 8, 10, 64, 0.9), (128, 8, 10, 64, 0.4), (128, 8, 1, 1, 0.5),
                                        (128, 8, 10, 10, 0.5), (128, 8, 33, 50, 0.7), (128, 8, 100, 100, 0.6), (128, 8, 10, 128, 0.6),
                                        (96, 16, 10, 64, 0.5), (96, 16, 10, 100, 0.7), (768, 32, 20, 64, 0.5)])
def test_filtered_walk_matches_the_oracle(D, M, k, ef, s):
    N = 20000 if D <= 128 else 6000
    e, cb, Q, labels, rng = make(N, D, M, 900 + ef + k + int(10 * s))
    g = O.Graph.from_state(e.get_graph(), M, 256)
    t = O.adc_table(Q, cb, 'euclidean')
    allow = labels[rng.random(N) < s]
    e.search(queries=Q[:4], k=k, ef=ef)                             # the first search after an insertion syncs the device graph
    n0 = e.launch_count
    l, d, st = e.search(queries=Q, k=k, ef=ef, filter_labels=allow, with_stats=True)        # fused table build
    assert e.launch_count - n0 == 3 and e.fallback_count == 0      # bitmap (2 kernels) + ONE walk kernel, no K1
    assert np.isin(l, allow).all()
    check_vs_oracle(l, d, st, g, t, k, ef, allow)
    # the literal tables= form (TMA staging), round 1's flagged walk and the bitmap walk: the same rows
    for name, opts in (('tables', {}), ('flagged', {'flagged_kernel': 1}), ('bitmap', {'force_general': 2})):
        for o, val in opts.items():
            e.set_option(o, val)
        l2, d2, st2 = e.search(tables=t, k=k, ef=ef, filter_labels=allow, with_stats=True)
        for o in opts:
            e.set_option(o, 0)
        if name == 'tables':
            assert np.array_equal(l, l2) and np.array_equal(bits(d), bits(d2)) and np.array_equal(st[:, :2], st2[:, :2])
        else:
            v = np.array(tie_aware_rows(l, d, l2, d2))
            assert (v == 'diff').sum() <= 1, name
    # and the scalar model of the two-list walk, where it flags nothing
    ml, md, mf, mh, mn, peak = O.two_list_walk(g, t, k, ef, filter_labels=allow, cap_n=256)
    assert (mf == k).all()
    v = np.array(tie_aware_rows(l, d, ml, md))
    assert (v == 'diff').sum() <= 1


@pytest.mark.parametrize('D,M,frac', [(128, 8, 0.1), (128, 8, 0.5), (96, 16, 0.2)])
def test_deletion_aware_walk_matches_the_oracle(D, M, frac):
    N = 20000
    e, cb, Q, labels, rng = make(N, D, M, 950 + int(100 * frac))
    dead = labels[rng.random(N) < frac]
    for x in dead:
        e.mark_deleted(int(x))
    g = O.Graph.from_state(e.get_graph(), M, 256)
    t = O.adc_table(Q, cb, 'euclidean')
    l, d, st = e.search(queries=Q, k=10, ef=64, with_stats=True)
    assert not np.isin(l, dead).any()
    check_vs_oracle(l, d, st, g, t, 10, 64, None)
    e.set_option('flagged_kernel', 1)
    l1, d1 = e.search(queries=Q, k=10, ef=64)
    e.set_option('flagged_kernel', 0)
    assert (np.array(tie_aware_rows(l, d, l1, d1)) == 'diff').sum() <= 1
    # a filter on top of deletions: delete marks are ignored under a filter (hnswalg.h:423-426)
    allow = labels[rng.random(N) < 0.5]
    lf, df, stf = e.search(queries=Q, k=10, ef=64, filter_labels=allow, with_stats=True)
    check_vs_oracle(lf, df, stf, g, t, 10, 64, allow)


def test_forced_route_without_filter_equals_the_plain_search():
    e, cb, Q, labels, rng = make(20000, 128, 8, 971)
    l0, d0, st0 = e.search(queries=Q, k=10, ef=64, with_stats=True)                 # hnsw_walk4
    e.set_option('force_general', 1)
    l1, d1, st1 = e.search(queries=Q, k=10, ef=64, with_stats=True)                 # hnsw_walk4f, everything passes
    e.set_option('force_general', 0)
    assert np.array_equal(l0, l1) and np.array_equal(bits(d0), bits(d1)) and np.array_equal(st0[:, :2], st1[:, :2])


def test_too_small_list_reruns_exactly_the_flagged_queries():
    N = 20000
    e, cb, Q, labels, rng = make(N, 128, 8, 981)
    g = O.Graph.from_state(e.get_graph(), 8, 256)
    t = O.adc_table(Q, cb, 'euclidean')
    allow = labels[rng.random(N) < 0.35]
    for en, cap in ((2, 64), (4, 128)):
        e.set_option('flagged_en', en)
        e.set_option('reset_counters', 0)
        l, d = e.search(queries=Q, k=10, ef=64, filter_labels=allow)
        e.set_option('flagged_en', 0)
        ml, md, mf, mh, mn, peak = O.two_list_walk(g, t, 10, 64, filter_labels=allow, cap_n=cap)
        assert e.fallback_queries == int((mf == -1).sum()), (en, e.fallback_queries, int((mf == -1).sum()))
        if en == 2:
            assert e.fallback_queries > 0
        check_vs_oracle(l, d, None, g, t, 10, 64, allow)


def test_too_few_admitted_results_raise_like_the_reference():
    e, cb, Q, labels, rng = make(5000, 128, 8, 985)
    with pytest.raises(RuntimeError, match='Cannot return the results in a contigious 2D array'):
        e.search(queries=Q, k=10, ef=64, filter_labels=labels[:3])
    with pytest.raises(RuntimeError, match='Cannot return the results in a contigious 2D array'):
        e.search(queries=Q, k=10, ef=64, filter_labels=np.zeros(0, dtype=np.uint64))


def test_streamed_filtered_search_equals_the_blocking_call():
    import torch
    N = 20000
    e, cb, Q, labels, rng = make(N, 128, 8, 990, nq=3000)
    allow = labels[rng.random(N) < 0.5]
    l, d = e.search(queries=Q, k=10, ef=64, filter_labels=allow)
    # host buffers (pinned), two batches in flight, the filter uploaded per batch on the batch's own lane
    Qp = torch.from_numpy(Q).pin_memory()
    fl = torch.from_numpy(allow.view(np.int64)).pin_memory()
    outs = [(torch.empty((3000, 10), dtype=torch.int64).pin_memory(), torch.empty((3000, 10), dtype=torch.float32).pin_memory())
            for _ in range(4)]
    n0, fb0 = e.launch_count, e.fallback_count
    tk = []
    for i in range(4):
        if i >= 2:
            e.search_wait(tk[i - 2])
        tk.append(e.search_submit(Qp.numpy(), outs[i][0].numpy().view(np.uint64), outs[i][1].numpy(), k=10, ef=64,
                                  filter_labels=fl.numpy().view(np.uint64)))
    e.search_wait(tk[2])
    e.search_wait(tk[3])
    if e.fallback_count == fb0:                                    # (a flagged query would add its re-run's launches)
        assert e.launch_count - n0 == 4 * 3                       # per batch: 2 bitmap kernels + ONE walk kernel
    for i in range(4):
        assert np.array_equal(outs[i][0].numpy().view(np.uint64), l) and np.array_equal(bits(outs[i][1].numpy()), bits(d))
    # device buffers, filter on the device
    Qd, fd = torch.from_numpy(Q).cuda(), torch.from_numpy(allow.view(np.int64)).cuda()
    od = [(torch.empty((3000, 10), dtype=torch.int64, device='cuda'), torch.empty((3000, 10), dtype=torch.float32, device='cuda'))
          for _ in range(2)]
    tk = [e.search_submit(Qd, od[i][0], od[i][1], k=10, ef=64, filter_labels=fd) for i in range(2)]
    for x in tk:
        e.search_wait(x)
    for i in range(2):
        assert np.array_equal(od[i][0].cpu().numpy().view(np.uint64), l) and np.array_equal(bits(od[i][1].cpu().numpy()), bits(d))
    # a list that is too small on purpose: the flagged queries are re-run inside search_wait
    e.set_option('flagged_en', 2)
    allow2 = labels[rng.random(N) < 0.35]
    lb, db = e.search(queries=Q, k=10, ef=64, filter_labels=allow2)
    e.set_option('reset_counters', 0)
    t0 = e.search_submit(Qp.numpy(), outs[0][0].numpy().view(np.uint64), outs[0][1].numpy(), k=10, ef=64, filter_labels=allow2)
    e.search_wait(t0)
    e.set_option('flagged_en', 0)
    assert e.fallback_queries > 0
    assert np.array_equal(outs[0][0].numpy().view(np.uint64), lb) and np.array_equal(bits(outs[0][1].numpy()), bits(db))
    # too few admitted nodes: the wait reports it
    t1 = e.search_submit(Qp.numpy(), outs[0][0].numpy().view(np.uint64), outs[0][1].numpy(), k=10, ef=64, filter_labels=labels[:3])
    with pytest.raises(RuntimeError, match='Cannot return the results in a contigious 2D array'):
        e.search_wait(t1)
    # deleted nodes are served by the streamed form too
    dead = labels[rng.random(N) < 0.2]
    for x in dead:
        e.mark_deleted(int(x))
    ld, dd = e.search(queries=Q, k=10, ef=64)
    t2 = e.search_submit(Qp.numpy(), outs[1][0].numpy().view(np.uint64), outs[1][1].numpy(), k=10, ef=64)
    e.search_wait(t2)
    assert not np.isin(ld, dead).any()
    assert np.array_equal(outs[1][0].numpy().view(np.uint64), ld) and np.array_equal(bits(outs[1][1].numpy()), bits(dd))

