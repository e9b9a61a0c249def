"""GPU: the reference-shaped Python surface (pq_bind / hnsw_bind.Index / PQCodec / HnswIndex / PQIndex /
AnnLite) -- these read like the reference's own tests (tests/test_pq_bind.py, tests/test_pq_index.py,
tests/test_hnsw_load_save.py, tests/test_dump.py) plus parity against the oracle."""
import pickle

import numpy as np
import pytest

import oracle as O
from helpers import bits, recall, tie_aware_rows

pytestmark = pytest.mark.gpu


def numpy_adc_table(query, n_subvectors, n_clusters, d_subvector, codebooks):
    dtable = np.empty((n_subvectors, n_clusters), dtype=np.float32)
    for m in range(n_subvectors):
        query_sub = query[m * d_subvector:(m + 1) * d_subvector]
        dtable[m, :] = np.linalg.norm(codebooks[m] - query_sub, axis=1) ** 2
    return dtable


# tests/test_pq_bind.py:36-59
def test_pq_adc_table_shape_and_values():
    from annlite_b200 import pq_bind
    rng = np.random.default_rng(0)
    n_data, n_dim, M, Ks = 1000, 128, 32, 256
    ds = n_dim // M
    X = rng.random((n_data, n_dim)).astype(np.float32)
    cb = rng.random((M, Ks, ds)).astype(np.float32)
    query = X[0]
    np_t = numpy_adc_table(query, M, Ks, ds, cb)
    t = np.asarray(pq_bind.precompute_adc_table(query, ds, Ks, cb))
    assert t.shape == np_t.shape
    np.testing.assert_array_almost_equal(t, np_t, decimal=5)
    assert np.array_equal(bits(t), bits(O.adc_table(query[None], cb)[0]))
    raw = pq_bind.batch_precompute_adc_table_ip(X[:7], ds, Ks, cb)
    ref = np.float32(1 / Ks) - O.adc_table(X[:7], cb, 'inner_product')   # not exact: compare to einsum
    want = np.einsum('mkd,nmd->nmk', cb, X[:7].reshape(7, M, ds))
    np.testing.assert_allclose(raw, want, rtol=1e-4, atol=1e-5)
    dists = pq_bind.dist_pqcodes_to_codebooks(np_t, rng.integers(0, Ks, (50, M)).astype(np.uint8))
    assert dists.shape == (50,)


# tests/test_pq_bind.py:62-75 and tests/test_pq_index.py:30-49
def test_codec_interfaces_agree(golden):
    from annlite_b200 import PQCodec, pq_bind
    codec = PQCodec(golden.M * golden.ds, n_subvectors=golden.M, n_clusters=golden.Ks, metric=golden.metric)
    with pytest.raises(AssertionError):
        codec._check_trained()
    codec.set_codebook(golden.cb)
    assert codec.is_trained and codec.get_subspace_splitting() == (golden.M, golden.Ks, golden.ds)
    q = golden.Q[:6]
    batched = codec.get_dist_mat(q)
    assert batched.dtype == np.float32 and batched.flags['C_CONTIGUOUS']
    stacked = np.stack([codec.get_dist_mat(q[i:i + 1])[0] for i in range(6)])
    assert np.allclose(batched, stacked) and np.array_equal(bits(batched), bits(stacked))
    t = codec.precompute_adc(q[0]).dtable
    assert np.array_equal(bits(t), bits(np.asarray(pq_bind.precompute_adc_table(q[0], golden.ds, golden.Ks, golden.cb))))
    # round trip through pickle (BaseCodec.dump/load)
    c2 = pickle.loads(pickle.dumps(codec))
    assert np.array_equal(c2.get_codebook(), golden.cb) and c2.is_trained
    codes = codec.encode(golden.X[:100] if golden.metric != 'cosine' else O.l2_normalize(golden.X[:100]).astype(np.float32))
    assert codes.dtype == codec.code_dtype and codes.shape == (100, golden.M)
    assert codec.decode(codes).shape == (100, golden.M * golden.ds)


def test_hnsw_bind_index_literal_call_shapes(golden):
    """Index used exactly the way HnswIndex uses the pybind11 class: codes + dtables in, (ids, dists) out."""
    from annlite_b200 import PQCodec
    from annlite_b200.hnsw_bind import Index
    codec = PQCodec(golden.M * golden.ds, n_subvectors=golden.M, n_clusters=golden.Ks, metric=golden.metric)
    codec.set_codebook(golden.cb)
    space = {'euclidean': 'l2', 'inner_product': 'ip', 'cosine': 'cosine'}[golden.metric]
    idx = Index(space=space, dim=golden.M * golden.ds)
    st = golden.state
    idx.init_index(max_elements=st['max_elements'], ef_construction=st['ef_construction'], M=st['M'], pq_codec=codec)
    assert idx.pq_enable and idx.element_count == 0 and idx.max_elements == st['max_elements']
    X = golden.X if golden.metric != 'cosine' else O.l2_normalize(golden.X).astype(np.float32)
    Tins = O.adc_table(X, golden.cb, golden.metric)          # what pre_process would hand over
    idx.set_num_threads(1)
    idx.add_items(golden.codes, ids=golden.labels, dtables=Tins)
    assert idx.element_count == len(golden.codes)
    idx.set_ef(golden.ef)
    tq = golden.query_tables_oracle()
    dummy_codes = np.zeros((len(tq), golden.M), dtype=codec.code_dtype)
    ids, dists = idx.knn_query(dummy_codes, k=golden.k, dtables=tq)
    assert ids.dtype == np.uint64 and dists.dtype == np.float32 and ids.shape == (len(tq), golden.k)
    v = tie_aware_rows(ids, dists, golden.knn_labels, golden.knn_dists)
    assert v.count('diff') <= (8 if golden.name == 'ties_k16' else 0)
    ids, dists = idx.knn_query_with_filter(dummy_codes, filters=golden.allow, k=golden.k, dtables=tq)
    assert np.isin(ids, golden.allow).all()
    with pytest.raises(RuntimeError, match='Cannot return the results'):
        idx.knn_query_with_filter(dummy_codes, filters=golden.allow[:2], k=golden.k, dtables=tq)
    with pytest.raises(RuntimeError, match='Label not found'):
        idx.mark_deleted(2 ** 40)
    # pickle round trip keeps graph + results (tests/test_hnsw_load_save.py spirit)
    idx2 = pickle.loads(pickle.dumps(idx))
    assert idx2.element_count == idx.element_count
    i2, d2 = idx2.knn_query(dummy_codes, k=golden.k, dtables=tq)
    i1, d1 = idx.knn_query(dummy_codes, k=golden.k, dtables=tq)
    assert np.array_equal(i1, i2) and np.array_equal(bits(d1), bits(d2))
    assert sorted(idx.get_ids_list()) == sorted(golden.labels.tolist())
    assert np.array_equal(np.asarray(idx.get_items(golden.labels[:4])), golden.codes[:4])
    with pytest.raises(ValueError, match='Initialization Error'):
        Index(space=space, dim=golden.M * golden.ds + golden.M).init_index(10, pq_codec=codec)


def test_hnsw_index_matches_reference_semantics(golden, tmp_path):
    """HnswIndex.search == reference HnswIndex.search: one query, sqrt for EUCLIDEAN, ids uint64."""
    from annlite_b200 import HnswIndex, Metric, PQCodec
    metric = Metric.from_string(golden.metric)
    codec = PQCodec(golden.M * golden.ds, n_subvectors=golden.M, n_clusters=golden.Ks, metric=metric)
    codec.set_codebook(golden.cb)
    st = golden.state
    h = HnswIndex(golden.M * golden.ds, metric=metric, pq_codec=codec, ef_search=golden.ef,
                  ef_construction=st['ef_construction'], max_connection=st['M'], initial_size=st['max_elements'])
    h._index._e.set_graph(st)   # adopt the reference-built graph for an apples-to-apples check
    h._index._cur_l = st['cur_element_count']
    assert h.size == st['cur_element_count'] and h.space_name in ('l2', 'ip', 'cosine')
    for qi in range(5):
        d, i = h.search(golden.Q[qi], limit=golden.k)
        ref_d = golden.knn_dists[qi]
        if golden.metric == 'euclidean':
            ref_d = np.sqrt(ref_d)
        assert i.shape == (golden.k,) and i.dtype == np.uint64
        if golden.metric == 'cosine':
            assert recall(i[None], golden.knn_labels[qi][None]) >= 0.9
        else:
            assert tie_aware_rows(i[None], d[None], golden.knn_labels[qi][None], ref_d[None])[0] != 'diff'
    d, i = h.search(golden.Q[0], limit=golden.k, indices=golden.allow)
    assert np.isin(i, golden.allow).all()
    # dump / load keeps size and answers (tests/test_hnsw_load_save.py:26-37)
    p = tmp_path / 'cell.hnsw'
    h.dump(p)
    h2 = HnswIndex(golden.M * golden.ds, metric=metric, pq_codec=codec, ef_search=golden.ef, index_file=p)
    assert h2.size == h.size
    d1, i1 = h.search_batch(golden.Q, limit=golden.k)
    d2, i2 = h2.search_batch(golden.Q, limit=golden.k)
    assert np.array_equal(i1, i2)
    # streamed form returns the same thing
    t1 = h.search_batch_submit(golden.Q[:32], limit=golden.k)
    t2 = h.search_batch_submit(golden.Q[32:], limit=golden.k)
    ds1, is1 = h.search_batch_wait(t1)
    ds2, is2 = h.search_batch_wait(t2)
    assert np.array_equal(np.vstack([is1, is2]), i1) and np.allclose(np.vstack([ds1, ds2]), d1, rtol=1e-6)
    # ... with a filter too (annb_search_submit_filtered; every code geometry, not only hnsw_walk4f's)
    df, jf = h.search_batch(golden.Q, limit=golden.k, indices=golden.allow)
    t3 = h.search_batch_submit(golden.Q[:32], limit=golden.k, indices=golden.allow)
    t4 = h.search_batch_submit(golden.Q[32:], limit=golden.k, indices=golden.allow)
    ds3, is3 = h.search_batch_wait(t3)
    ds4, is4 = h.search_batch_wait(t4)
    assert np.array_equal(np.vstack([is3, is4]), jf) and np.allclose(np.vstack([ds3, ds4]), df, rtol=1e-6)
    assert np.isin(jf, golden.allow).all()
    h.delete([int(golden.labels[0])])
    with pytest.raises(RuntimeError, match='update operation is not allowed'):
        h.update_with_ids(golden.X[:1], [0])
    # after a deletion the streamed form still answers (deletion-aware walk), and never with the deleted id
    t5 = h.search_batch_submit(golden.Q, limit=golden.k)
    ds5, is5 = h.search_batch_wait(t5)
    d5, i5 = h.search_batch(golden.Q, limit=golden.k)
    assert np.array_equal(is5, i5) and not (is5 == golden.labels[0]).any()


def test_pq_index_linear_scan(golden):
    from annlite_b200 import PQCodec, PQIndex
    if golden.metric != 'euclidean':
        pytest.skip('PQIndex tables are squared-L2 (pq.py:200-224)')
    codec = PQCodec(golden.M * golden.ds, n_subvectors=golden.M, n_clusters=golden.Ks)
    codec.set_codebook(golden.cb)
    n = len(golden.X)
    idx = PQIndex(golden.M * golden.ds, codec, initial_size=n)   # sized exactly (SURVEY 3.3 caveat)
    idx.add_with_ids(golden.X, list(range(n)))
    assert idx.size == n
    codes = codec.encode(golden.X)
    t = O.adc_table(golden.Q[:8], golden.cb)
    oi, od = O.scan_topk(t, codes, 10)
    for qi in range(8):
        d, i = idx.search(golden.Q[qi], limit=10)
        assert np.array_equal(i, oi[qi]) and np.array_equal(bits(d), bits(od[qi]))
    sub = np.arange(0, n, 3)
    d, i = idx.search(golden.Q[0], limit=5, indices=sub)
    assert np.isin(i, sub).all()


def test_annlite_index_search_dump_restore(tmp_path):
    """tests/test_dump.py:23-42 spirit: same top-10 after reload."""
    from annlite_b200 import AnnLite
    rng = np.random.default_rng(7)
    N, D = 3000, 32
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((20, D)).astype(np.float32)
    a = AnnLite(D, metric='euclidean', n_subvectors=4, n_clusters=64, data_path=tmp_path / 'ws', ef_search=64,
                initial_size=N)
    with pytest.raises(RuntimeError, match='not trained'):
        a.index(X)
    a.train(X[:1000], iter=5, random_state=0)
    assert a.is_trained
    off = a.index(X)
    assert len(off) == N and a.index_size == N and a.stat['total_docs'] == N
    d, i = a.search(Q, limit=10)
    assert d.shape == (20, 10) and (np.diff(d, axis=1) >= 0).all()
    # parity of the whole stack against the oracle on the graph the product built
    e = a._index._index._e
    g = O.Graph.from_state(e.get_graph(), 4, 64)
    t = O.adc_table(Q, a._pq_codec.codebooks)
    ol, od, _ = O.hnsw_search(g, t, 10, 64)
    assert tie_aware_rows(i, d, ol, np.sqrt(od)).count('diff') == 0
    mask = np.zeros(N, bool)
    mask[::2] = True
    d2, i2 = a.search_numpy(Q, filter=mask, limit=5)
    assert (i2 % 2 == 0).all()
    a.dump()
    b = AnnLite(D, metric='euclidean', n_subvectors=4, n_clusters=64, data_path=tmp_path / 'ws', ef_search=64,
                initial_size=N)
    assert b.is_trained and b.index_size == N
    d3, i3 = b.search(Q, limit=10)
    assert np.array_equal(i, i3)
    a.delete([int(i[0, 0])])
    d4, i4 = a.search(Q[:1], limit=10)
    assert int(i[0, 0]) not in i4
    # update: move a stored vector onto a query -> it becomes that query's nearest neighbour
    victim = int(i[1, 5])
    a.update(Q[1:2], ids=[victim])
    assert a.index_size == N
    d5, i5 = a.search(Q[1:2], limit=10)
    assert victim in i5[0][:3]


def test_merge_kernel_matches_host_rule():
    torch = pytest.importorskip('torch')
    from annlite_b200.engine import Engine
    from annlite_b200.sharded import merge_topk_host
    rng = np.random.default_rng(1)
    e = Engine(8, 2, 16)
    for (G, B, k) in [(2, 33, 10), (8, 100, 100), (4, 7, 1)]:
        d = np.sort(rng.random((G, B, k)).astype(np.float32), axis=2)
        l = rng.permutation(G * B * k).astype(np.uint64).reshape(G, B, k)
        d[1, 0, :] = d[0, 0, :]
        if k > 2:
            l[G - 1, 1, 2:] = np.iinfo(np.uint64).max
            d[G - 1, 1, 2:] = np.inf
        tl = torch.from_numpy(l.view(np.int64)).cuda()
        td = torch.from_numpy(d).cuda()
        ol = torch.empty((B, k), dtype=torch.int64, device='cuda')
        od = torch.empty((B, k), dtype=torch.float32, device='cuda')
        e.merge_topk(tl, td, ol, od)
        e.sync()
        el, ed = merge_topk_host(l, d, k)
        assert np.array_equal(ol.cpu().numpy().view(np.uint64), el) and np.array_equal(od.cpu().numpy(), ed)


def test_annlite_explicit_then_implicit_ids_do_not_collide(tmp_path):
    """Offsets handed out after a caller-supplied id range continue past it (an id reused would be an in-place
    update of the earlier vector: hnswalg.h:1119-1131), also after restore() of sparse labels; an empty filter
    returns empty rows and an oversized limit is refused with a clear message."""
    from annlite_b200 import AnnLite
    rng = np.random.default_rng(8)
    D = 32
    X = rng.standard_normal((400, D)).astype(np.float32)
    a = AnnLite(D, metric='euclidean', n_subvectors=4, n_clusters=64, data_path=tmp_path / 'ws2', ef_search=32, initial_size=2000)
    a.train(X, iter=5, random_state=0)
    a.index(X[:100], ids=np.arange(100, 200))
    off = a.index(X[100:200])
    assert off.min() >= 200 and a.index_size == 200
    a.dump()
    b = AnnLite(D, metric='euclidean', n_subvectors=4, n_clusters=64, data_path=tmp_path / 'ws2', ef_search=32, initial_size=2000)
    off2 = b.index(X[200:300])
    assert off2.min() >= 300 and b.index_size == 300
    d, i = b.search_numpy(X[:3], filter=np.zeros(0, dtype=np.uint64), limit=5)
    assert d.shape == (3, 0) and i.shape == (3, 0)
    with pytest.raises(ValueError, match='ANNB_MAX_EF'):
        b.search_numpy(X[:3], limit=600)


def test_general_walk_scratch_grows_when_almost_everything_is_deleted():
    """the literal (bitmap) walk must not fail when it has to visit a large part of the graph -- its
    per-query scratch (visited log, candidate bag) grows instead.  With all but 150 nodes deleted the reference's
    searchBaseLayerST<has_deletions> keeps expanding until it holds ef live nodes (hnswalg.h:270 waits for size == ef)."""
    from annlite_b200.engine import Engine
    rng = np.random.default_rng(9)
    N, D, M = 60_000, 32, 8
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((8, D)).astype(np.float32)
    cb = np.stack([X[rng.choice(N, 256, replace=False), m * 4:(m + 1) * 4] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, 256, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(N, M=16, ef_construction=64)
    e.add_items(X, np.arange(N, dtype=np.uint64))
    keep = set(rng.choice(N, 150, replace=False).tolist())
    for i in range(N):
        if i not in keep:
            e.mark_deleted(i)
    e.set_option('force_general', 2)
    l, d, st = e.search(queries=Q, k=5, ef=32, with_stats=True)
    assert np.isin(l, np.fromiter(keep, dtype=np.uint64)).all()
    assert st[:, 2].max() > 32768                     # more evaluations than the initial visited log holds
    g = O.Graph.from_state(e.get_graph(), M, 256)
    ol, od, found = O.hnsw_search(g, O.adc_table(Q, cb), 5, 32)
    assert tie_aware_rows(l, d, ol, od).count('diff') == 0
