"""GPU: edge cases of the search path -- tiny / empty indexes, k vs ef vs n, huge sparse labels, empty and
degenerate filters, incremental growth with device re-sync, argument errors."""
import numpy as np
import pytest

import oracle as O
from annlite_b200 import _lib as L
from annlite_b200.engine import Engine
from helpers import bits, tie_aware_rows

pytestmark = pytest.mark.gpu


def small(n, D=16, M=4, Ks=16, seed=0, labels=None):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((max(n, Ks), D)).astype(np.float32)
    ds = D // M
    cb = np.stack([X[rng.choice(len(X), Ks, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, Ks, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(max(n, 1) + 8, M=4, ef_construction=20)
    lab = np.arange(n, dtype=np.uint64) if labels is None else labels
    if n:
        e.add_items(X[:n], lab, num_threads=1)
    Q = rng.standard_normal((9, D)).astype(np.float32)
    return e, cb, X[:n], Q, lab


def oracle_same(e, cb, Q, k, ef, flt=None, M=4, Ks=16):
    g = O.Graph.from_state(e.get_graph(), M, Ks)
    t = O.adc_table(Q, cb)
    ol, od, found = O.hnsw_search(g, t, k, ef, filter_labels=flt)
    return ol, od, found


def test_empty_index_raises_like_the_reference():
    e, cb, X, Q, lab = small(0)
    with pytest.raises(RuntimeError, match='Cannot return the results in a contigious 2D array'):
        e.search(queries=Q, k=1, ef=10)
    assert e.element_count == 0
    l, d = e.search(queries=Q[:0], k=1, ef=10)      # zero queries is fine
    assert l.shape == (0, 1)


@pytest.mark.parametrize('n', [1, 2, 3, 5, 33])
def test_tiny_indexes(n):
    e, cb, X, Q, lab = small(n)
    k = min(n, 3)
    l, d = e.search(queries=Q, k=k, ef=10)
    ol, od, found = oracle_same(e, cb, Q, k, 10)
    assert (found == k).all()
    assert tie_aware_rows(l, d, ol, od).count('diff') == 0
    if n < 4:
        with pytest.raises(RuntimeError, match='Cannot return the results'):
            e.search(queries=Q, k=n + 1, ef=10)      # fewer than k reachable results


def test_k_larger_than_ef_uses_max():
    e, cb, X, Q, lab = small(400)
    l, d = e.search(queries=Q, k=40, ef=5)           # searchKnn uses max(ef, k)  (hnswalg.h:1279)
    ol, od, _ = oracle_same(e, cb, Q, 40, 5)
    assert tie_aware_rows(l, d, ol, od).count('diff') == 0
    with pytest.raises(L.AnnbError) as ei:
        e.search(queries=Q, k=10, ef=L.MAX_EF + 1)
    assert ei.value.code == L.ELIMIT


def test_huge_sparse_labels_and_filters():
    rng = np.random.default_rng(3)
    labels = np.unique(rng.integers(1, 2 ** 62, 600, dtype=np.uint64))[:500]
    rng.shuffle(labels)
    e, cb, X, Q, lab = small(500, labels=labels)
    l, d = e.search(queries=Q, k=5, ef=32)
    assert np.isin(l, labels).all()
    ol, od, _ = oracle_same(e, cb, Q, 5, 32)
    assert tie_aware_rows(l, d, ol, od).count('diff') == 0
    allow = labels[::2]
    l, d = e.search(queries=Q, k=5, ef=32, filter_labels=allow)   # host-side label resolution path
    assert np.isin(l, allow).all()
    ol, od, _ = oracle_same(e, cb, Q, 5, 32, flt=allow)
    assert tie_aware_rows(l, d, ol, od).count('diff') == 0
    # labels that are not in the index are ignored, as a fuse filter built from them would be
    l2, d2 = e.search(queries=Q, k=5, ef=32, filter_labels=np.concatenate([allow, np.array([7, 9], dtype=np.uint64)]))
    assert np.array_equal(l, l2)


def test_degenerate_filters():
    e, cb, X, Q, lab = small(300)
    with pytest.raises(RuntimeError, match='Cannot return the results'):
        e.search(queries=Q, k=3, ef=16, filter_labels=np.zeros(0, dtype=np.uint64))
    one = lab[[17]]
    ol, od, found = oracle_same(e, cb, Q, 1, 16, flt=one)
    try:
        l, d = e.search(queries=Q, k=1, ef=16, filter_labels=one)
        assert (found == 1).all() and (l == 17).all()
    except RuntimeError:
        assert (found < 1).any()       # the reference's filtered walk can stop early too (no size==ef guard)


def test_incremental_growth_resyncs_device_graph():
    rng = np.random.default_rng(5)
    D, M, Ks = 16, 4, 16
    X = rng.standard_normal((600, D)).astype(np.float32)
    cb = np.stack([X[rng.choice(600, Ks, replace=False), m * 4:(m + 1) * 4] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, Ks, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(100, M=8, ef_construction=40)
    Q = rng.standard_normal((16, D)).astype(np.float32)
    done = 0
    for step in (50, 50, 200, 300):
        if done + step > 100 and e.graph_info()['max_elements'] < done + step:
            e.resize_index(done + step)
        e.add_items(X[done:done + step], np.arange(done, done + step, dtype=np.uint64), num_threads=1)
        done += step
        l, d = e.search(queries=Q, k=5, ef=32)
        assert (l < done).all()
        ol, od, _ = oracle_same(e, cb, Q, 5, 32)
        assert tie_aware_rows(l, d, ol, od).count('diff') == 0
    # the incremental single-threaded graph equals a one-shot single-threaded build
    e2 = Engine(D, M, Ks, 'euclidean')
    e2.set_codebook(cb)
    e2.init_graph(600, M=8, ef_construction=40)
    e2.add_items(X, np.arange(600, dtype=np.uint64), num_threads=1)
    a, b = e.get_graph(), e2.get_graph()
    assert np.array_equal(a['data_level0'], b['data_level0']) and np.array_equal(a['link_lists'], b['link_lists'])


def test_argument_errors():
    e, cb, X, Q, lab = small(50)
    with pytest.raises(L.AnnbError):
        e._lib.annb_search(e._h, None, None, 0, 1, 0, 1, 10, None, 0, 0, Q.ctypes.data, Q.ctypes.data, 0, None) and L.check(-1)
    rc = e._lib.annb_search(e._h, None, None, 0, 1, 0, 1, 10, None, 0, 0, Q.ctypes.data, Q.ctypes.data, 0, None)
    assert rc == L.EINVAL and b'exactly one of' in e._lib.annb_last_error()
    e.add_items(X[:1], lab[:1])          # re-adding a stored label is an in-place update, not an error
    assert e.element_count == 50
    e2 = Engine(16, 4, 16)
    with pytest.raises(L.AnnbError, match='train the PQ'):
        e2.adc_table(Q)
    with pytest.raises(ValueError, match='Initialization Error'):
        Engine(18, 4, 16)


def test_bruteforce_subset_scan_is_exact_over_the_subset():
    """exact ADC over a small admissible id set (opt-in route for very selective filters)."""
    e, cb, X, Q, lab = small(800, seed=21, labels=np.arange(800, dtype=np.uint64) * 5 + 2)
    sub = lab[::37]                                      # 22 admissible labels
    l, d = e.scan_subset(Q, sub, k=5)
    g = O.Graph.from_state(e.get_graph(), 4, 16)
    codes, glab = g.codes(), g.labels()
    pos = {int(v): i for i, v in enumerate(glab)}
    t = O.adc_table(Q, cb)
    for b in range(len(Q)):
        dd = O.scan(t[b], codes[[pos[int(x)] for x in sub]])
        order = np.lexsort((np.arange(len(sub)), dd))[:5]
        assert np.array_equal(l[b], sub[order]) and np.array_equal(bits(d[b]), bits(dd[order]))
    # unknown labels are skipped, too few candidates raise like the walk does
    l2, d2 = e.scan_subset(Q, np.concatenate([sub, np.array([1, 3], dtype=np.uint64)]), k=5)
    assert np.array_equal(l, l2)
    with pytest.raises(RuntimeError, match='Cannot return the results'):
        e.scan_subset(Q, sub[:3], k=5)
    # the flat code matrix of the handle (annb_set_codes) is left alone
    e.set_codes(codes)
    before = e.scan_topk(queries=Q, k=3)[0]
    e.scan_subset(Q, sub, k=5)
    assert np.array_equal(before, e.scan_topk(queries=Q, k=3)[0])


def test_hnsw_index_opt_in_bruteforce_for_selective_filters():
    from annlite_b200 import HnswIndex, Metric, PQCodec
    rng = np.random.default_rng(31)
    N, D = 3000, 16
    X = rng.standard_normal((N, D)).astype(np.float32)
    codec = PQCodec(D, n_subvectors=4, n_clusters=16, metric=Metric.EUCLIDEAN)
    codec.set_codebook(np.stack([X[rng.choice(N, 16, replace=False), m * 4:(m + 1) * 4] for m in range(4)]))
    ref = HnswIndex(D, metric=Metric.EUCLIDEAN, pq_codec=codec, ef_search=32, initial_size=N)
    bf = HnswIndex(D, metric=Metric.EUCLIDEAN, pq_codec=codec, ef_search=32, initial_size=N, bruteforce_filter_below=64)
    for h in (ref, bf):
        h.add_with_ids(X, list(range(N)), num_threads=1)
    q = rng.standard_normal(D).astype(np.float32)
    few = np.sort(rng.choice(N, 12, replace=False)).astype(np.uint64)
    d, i = bf.search(q, limit=5, indices=few)                    # exact over the 12 admissible ids
    assert np.isin(i, few).all() and (np.diff(d) >= 0).all()
    codes = codec.encode(X[few.astype(np.int64)])
    exact = np.sqrt(O.scan(O.adc_table(q[None], codec.codebooks)[0], codes))
    assert np.allclose(np.sort(exact)[:5], d, rtol=1e-6)
    try:                                                         # the reference-shaped walk may find fewer / others
        d2, i2 = ref.search(q, limit=5, indices=few)
        assert np.isin(i2, few).all() and d2[0] >= d[0] - 1e-6
    except RuntimeError as ex:
        assert 'Cannot return the results' in str(ex)
    many = np.arange(0, N, 2, dtype=np.uint64)                   # above the threshold: both take the graph walk
    da, ia = bf.search(q, limit=5, indices=many)
    db, ib = ref.search(q, limit=5, indices=many)
    assert np.array_equal(ia, ib)


def test_small_insertions_patch_the_device_graph_instead_of_rederiving_it():
    """Interleaved index / search (the normal AnnLite usage): after the first full derivation of the walk layout, a
    small host insertion uploads only the level-0 records it rewrote (+ the upper levels when one of them moved);
    every search in between must still equal the oracle on the host graph, threaded insertions and an in-place
    update (which falls back to the full derivation) included."""
    rng = np.random.default_rng(15)
    N0, D, M = 30_000, 32, 8
    X = rng.standard_normal((N0 + 4000, D)).astype(np.float32)
    Q = rng.standard_normal((64, D)).astype(np.float32)
    cb = np.stack([X[rng.choice(N0, 256, replace=False), m * 4:(m + 1) * 4] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, 256, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(N0 + 4000, M=16, ef_construction=64)
    e.add_items(X[:N0], np.arange(N0, dtype=np.uint64))
    t = O.adc_table(Q, cb)

    def check():
        l, d = e.search(queries=Q, k=10, ef=64)
        g = O.Graph.from_state(e.get_graph(), M, 256)
        ol, od, found, ties = O.hnsw_search(g, t, 10, 64, with_ties=True)
        v = np.array(tie_aware_rows(l, d, ol, od))
        assert (v[ties == 0] == 'exact').all() and (v == 'diff').sum() <= 1

    check()
    full0, patch0 = e.sync_counts
    done = N0
    for step, threads in [(1, 1), (7, 1), (100, 1), (500, 4), (1000, 8), (3, 1)]:
        e.add_items(X[done:done + step], np.arange(done, done + step, dtype=np.uint64), num_threads=threads)
        done += step
        check()
    full1, patch1 = e.sync_counts
    # six insertions, six updates of the device copy; the 1000-row one rewrites more than a quarter of this 30k-node
    # graph's records and is allowed to re-derive everything instead
    assert (full1 - full0) + (patch1 - patch0) == 6 and patch1 - patch0 >= 5 and full1 - full0 <= 1
    # two insertions before one search: one patch covers both
    e.add_items(X[done:done + 10], np.arange(done, done + 10, dtype=np.uint64), num_threads=1)
    e.add_items(X[done + 10:done + 20], np.arange(done + 10, done + 20, dtype=np.uint64), num_threads=1)
    done += 20
    check()
    assert e.sync_counts == (full1, patch1 + 1)
    # re-adding a stored label is updatePoint: rewrites whole neighbourhoods -> full derivation, still correct
    e.add_items(X[done:done + 1], np.array([5], dtype=np.uint64), num_threads=1)
    check()
    assert e.sync_counts[0] == full1 + 1
    # deletions after a patch
    e.mark_deleted(7)
    l, d = e.search(queries=Q, k=10, ef=64)
    assert not (l == 7).any()
