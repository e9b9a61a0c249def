"""GPU parity (through the C ABI): CUDA kernels vs the oracle and vs what the compiled reference
returned (tests/golden).  Integer/index results must be identical (tie-aware where the reference
itself is order-dependent); fp32 tables/distances bit-identical for L2/IP; 1e-4 relative for the
cosine path whose l2_normalize step cannot be bit-reproduced (numpy einsum order)."""
import numpy as np
import pytest

import oracle as O
from annlite_b200.engine import Engine
from helpers import bits, recall, tie_aware_rows

pytestmark = pytest.mark.gpu

NORM = {'euclidean': 0, 'inner_product': 0, 'cosine': 2}


def engine(fx, with_graph=True, deleted=False):
    e = Engine(fx.M * fx.ds, fx.M, fx.Ks, fx.metric, device=0)
    e.set_codebook(fx.cb)
    if with_graph:
        e.set_graph(fx.state)
        if deleted:
            for l in fx.deleted:
                e.mark_deleted(int(l))
    return e


def prenorm(fx):
    Q = fx.Q
    if fx.metric == 'cosine':
        Q = O.l2_normalize(O.l2_normalize(Q).astype(np.float32)).astype(np.float32)
    return Q


# ---- K1 -------------------------------------------------------------------------------------
def test_k1_tables_bit_exact_vs_reference(golden):
    e = engine(golden, with_graph=False)
    t = e.adc_table(prenorm(golden)[:8], normalize=0)
    assert np.array_equal(bits(t), bits(golden.tables))


def test_k1_with_device_normalise(golden):
    e = engine(golden, with_graph=False)
    t = e.adc_table(golden.Q, normalize=NORM[golden.metric])
    ref = golden.query_tables_oracle()
    if golden.metric == 'cosine':
        assert np.allclose(t, ref, rtol=1e-5, atol=1e-6)
    else:
        assert np.array_equal(bits(t), bits(ref))


def test_k1_odd_shapes():
    rng = np.random.default_rng(5)
    for (M, Ks, ds, metric) in [(3, 7, 5, 'euclidean'), (2, 513, 3, 'inner_product'), (16, 256, 6, 'euclidean'),
                                (32, 256, 24, 'euclidean'), (1, 1, 1, 'inner_product'), (5, 70, 40, 'euclidean')]:
        cb = rng.standard_normal((M, Ks, ds)).astype(np.float32)
        q = rng.standard_normal((37, M * ds)).astype(np.float32)
        e = Engine(M * ds, M, Ks, metric)
        e.set_codebook(cb)
        assert np.array_equal(bits(e.adc_table(q)), bits(O.adc_table(q, cb, metric))), (M, Ks, ds, metric)
    e = Engine(8, 2, 4, 'euclidean')
    e.set_codebook(rng.standard_normal((2, 4, 4)).astype(np.float32))
    assert e.adc_table(np.zeros((0, 8), np.float32)).shape == (0, 2, 4)


# ---- K2 -------------------------------------------------------------------------------------
def test_k2_scan_bit_exact(golden):
    e = engine(golden, with_graph=False)
    e.set_codes(golden.codes)
    t = golden.query_tables_oracle()
    assert np.array_equal(bits(e.scan(t[0])), bits(golden.scan_d))


@pytest.mark.parametrize('k', [1, 10, 33, 100])
def test_k2_scan_topk_ids_exact(golden, k):
    e = engine(golden, with_graph=False)
    e.set_codes(golden.codes)
    t = golden.query_tables_oracle()
    ids, d = e.scan_topk(tables=t, k=k)
    rid, rd = O.scan_topk(t, golden.codes, k)
    assert np.array_equal(ids, rid)          # both order ties by (dist, row)
    assert np.array_equal(bits(d), bits(rd))


def test_k2_fused_queries_path(golden):
    if golden.metric != 'euclidean':
        pytest.skip('PQIndex.search builds the L2 table only (pq.py:200-224)')
    e = engine(golden, with_graph=False)
    e.set_codes(golden.codes)
    ids, d = e.scan_topk(queries=golden.Q, k=10)
    rid, rd = O.scan_topk(O.adc_table(golden.Q, golden.cb, 'euclidean'), golden.codes, 10)
    assert np.array_equal(ids, rid) and np.array_equal(bits(d), bits(rd))


def test_k2_k_larger_than_n():
    rng = np.random.default_rng(3)
    cb = rng.standard_normal((4, 16, 2)).astype(np.float32)
    codes = rng.integers(0, 16, (5, 4)).astype(np.uint8)
    e = Engine(8, 4, 16)
    e.set_codebook(cb)
    e.set_codes(codes)
    q = rng.standard_normal((3, 8)).astype(np.float32)
    ids, d = e.scan_topk(queries=q, k=8)
    rid, rd = O.scan_topk(O.adc_table(q, cb), codes, 8)
    assert np.array_equal(ids, rid)
    assert np.array_equal(ids[:, 5:], -np.ones((3, 3), np.int64)) and np.isinf(d[:, 5:]).all()


# ---- K3 -------------------------------------------------------------------------------------
def _check(fx, l, d, ref_l, ref_d, allow_diff=0):
    v = tie_aware_rows(l, d, ref_l, ref_d)
    ndiff = v.count('diff')
    assert ndiff <= allow_diff, f'{fx.name}: {ndiff} rows differ beyond ties ({v.count("tie")} tie rows)'
    return v


@pytest.mark.parametrize('general', [0, 1, 2])   # 0 fast walk, 1 flagged single-list walk, 2 bitmap walk
def test_k3_knn_matches_reference(golden, general):
    e = engine(golden)
    e.set_option('force_general', general)
    t = golden.query_tables_oracle()
    l, d, st = e.search(tables=t, k=golden.k, ef=golden.ef, with_stats=True)
    allow = 8 if golden.name == 'ties_k16' else 0   # exact-tie-heavy data: expansion order may diverge
    v = _check(golden, l, d, golden.knn_labels, golden.knn_dists, allow)
    g = golden.oracle_graph()
    _, _, _, (hops, nbrs, evals), ties = O.hnsw_search(g, t, golden.k, golden.ef, with_counts=True, with_ties=True)
    same = np.array([x == 'exact' for x in v])
    assert same[ties == 0].all()          # no tie met by the reference walk => no freedom at all
    if general != 2:
        same &= ties == 0                 # a tie the reference resolved by heap order can cost/save a hop
    assert np.array_equal(st[same, 0], hops[same]) and np.array_equal(st[same, 1], nbrs[same])
    if general == 2:   # exact visited set => the same number of distance evaluations as the reference
        assert np.array_equal(st[same, 2], evals[same])


def test_k3_fused_table_build(golden):
    e = engine(golden)
    l, d = e.search(queries=golden.Q, k=golden.k, ef=golden.ef, normalize=NORM[golden.metric])
    if golden.metric == 'cosine':
        assert recall(l, golden.knn_labels) >= 0.995
        m = l == golden.knn_labels
        assert np.allclose(d[m], golden.knn_dists[m], rtol=1e-4, atol=1e-6)
    else:
        _check(golden, l, d, golden.knn_labels, golden.knn_dists, 8 if golden.name == 'ties_k16' else 0)


@pytest.mark.parametrize('mode', ['auto', 'bitmap', 'tiny_list', 'partial'])
def test_k3_filtered_matches_reference(golden, mode):
    e = engine(golden)
    if mode == 'bitmap':
        e.set_option('force_general', 2)
    if mode == 'tiny_list':          # a list too small for a 50 % filter: queries overflow and are re-run on
        e.set_option('flagged_epl', 2)                                 # the bitmap walk
    if mode == 'partial':            # 128 slots: some queries fit, some do not -- only those are re-run
        e.set_option('flagged_epl', 4)
    t = golden.query_tables_oracle()
    l, d = e.search(tables=t, k=golden.k, ef=golden.ef, filter_labels=golden.allow)
    if mode == 'tiny_list' and golden.ef >= 50:   # 64 slots cannot hold ~2*ef candidates
        assert e.fallback_count >= 1
    if mode in ('tiny_list', 'partial') and 32 * (2 if mode == 'tiny_list' else 4) >= golden.ef:
        # the scalar model of the flagged walk says exactly which queries outgrow the list: only those are redone
        *_, mf, _, _, _ = O.flagged_walk(golden.oracle_graph(), t, golden.k, golden.ef, filter_labels=golden.allow,
                                         cap=64 if mode == 'tiny_list' else 128)
        assert e.fallback_queries == int((mf == -1).sum())
    assert np.isin(l, golden.allow).all()
    _check(golden, l, d, golden.flt_labels, golden.flt_dists, 8 if golden.name == 'ties_k16' else 0)


@pytest.mark.parametrize('force', [0, 2])
def test_k3_deleted_matches_reference(golden, force):
    e = engine(golden, deleted=True)
    e.set_option('force_general', force)
    t = golden.query_tables_oracle()
    l, d = e.search(tables=t, k=golden.k, ef=golden.ef)
    assert not np.isin(l, golden.deleted).any()
    _check(golden, l, d, golden.del_labels, golden.del_dists, 8 if golden.name == 'ties_k16' else 0)


@pytest.mark.parametrize('ef,k', [(10, 10), (33, 7), (100, 100), (200, 50), (300, 10)])
def test_k3_ef_k_sweep_vs_oracle(golden, ef, k):
    e = engine(golden)
    t = golden.query_tables_oracle()
    l, d = e.search(tables=t, k=k, ef=ef)
    rl, rd, found = O.hnsw_search(golden.oracle_graph(), t, k, ef)
    assert (found == k).all()
    _check(golden, l, d, rl, rd, 10 if golden.name == 'ties_k16' else 0)


def test_k3_too_few_results_raises(golden):
    e = engine(golden)
    t = golden.query_tables_oracle()
    few = golden.allow[:3]
    with pytest.raises(RuntimeError, match='Cannot return the results in a contigious 2D array'):
        e.search(tables=t, k=golden.k, ef=golden.ef, filter_labels=few)


def test_k3_device_buffers_roundtrip(golden):
    torch = pytest.importorskip('torch')
    e = engine(golden)
    t = torch.from_numpy(golden.query_tables_oracle()).cuda()
    B = t.shape[0]
    ol = torch.empty((B, golden.k), dtype=torch.int64, device='cuda')  # bit pattern of uint64 labels
    od = torch.empty((B, golden.k), dtype=torch.float32, device='cuda')
    e.search(tables=t, k=golden.k, ef=golden.ef, out_labels=ol, out_dists=od)
    e.sync()
    _check(golden, ol.cpu().numpy().view(np.uint64), od.cpu().numpy(), golden.knn_labels, golden.knn_dists,
           8 if golden.name == 'ties_k16' else 0)


# ---- build through the GPU table feed --------------------------------------------------------
def test_add_items_gpu_tables_single_thread_graph_identical(golden):
    e = engine(golden, with_graph=False)
    st = golden.state
    e.init_graph(st['max_elements'], M=st['M'], ef_construction=st['ef_construction'])
    X = golden.X
    if golden.metric == 'cosine':
        X = O.l2_normalize(X).astype(np.float32)   # pre_process; the library re-normalises for the tables
    e.add_items(X, golden.labels, codes=golden.codes, num_threads=1)
    got = e.get_graph()
    if golden.metric == 'cosine':   # device normalise is tolerance-level => graph may differ slightly
        assert got['cur_element_count'] == st['cur_element_count']
    else:
        assert np.array_equal(got['data_level0'], st['data_level0'])
        assert np.array_equal(got['link_lists'], st['link_lists'])
    # and it is searchable end to end
    l, d = e.search(queries=golden.Q, k=golden.k, ef=golden.ef, normalize=NORM[golden.metric])
    assert recall(l, golden.knn_labels) >= (0.9 if golden.metric == 'cosine' else 0.97)


def test_encode_matches_exact_argmin(golden):
    e = engine(golden, with_graph=False)
    X = golden.X
    if golden.metric == 'cosine':
        X = O.l2_normalize(X).astype(np.float32)
    c = e.encode(X)
    ref = O.encode(X, golden.cb)
    mism = float((c != ref).mean())
    assert mism <= 2e-3, mism       # fp32 vs fp64 evaluation differ on near-ties only
    vs_ref = float((c != golden.codes).mean())   # golden.codes = scipy vq inside the reference
    assert vs_ref <= 5e-3, vs_ref


# ---- K2, query-tiled kernel (scan_topk_tiled_kernel: lanes = queries, interleaved 16-query table tile) ---------
@pytest.mark.parametrize('k', [1, 7, 10, 16])
@pytest.mark.parametrize('N,B,Ks,ties', [(100_000, 200, 256, False), (40_000, 70, 256, True), (33_000, 64, 200, False),
                                         (2_500, 37, 256, False)])
def test_k2_tiled_scan_topk_ids_exact(N, B, Ks, ties, k):
    """ids and distance bits equal to the oracle's (dist, row) order, and to the round-1 kernel's; `ties` = a
    codebook of small integers, so many rows have exactly equal distances; the last shape forces the tiled
    kernel on an input the dispatcher would leave to the round-1 kernel (ragged tile, rows not a multiple of 32)."""
    rng = np.random.default_rng(N + B + k)
    cb = rng.standard_normal((8, Ks, 4)).astype(np.float32)
    if ties:
        cb = np.round(cb)
    codes = rng.integers(0, Ks, (N, 8)).astype(np.uint8)
    q = rng.standard_normal((B, 32)).astype(np.float32)
    if ties:
        q = np.round(q)
    e = Engine(32, 8, Ks)
    e.set_codebook(cb)
    e.set_codes(codes)
    t = O.adc_table(q, cb)
    e.set_option('scan_kernel', 2)
    n0 = e.launch_count
    ids, d = e.scan_topk(tables=t, k=k)
    assert e.launch_count - n0 == 2            # the tiled scan + the merge of the per-warp partial lists
    e.set_option('scan_kernel', 1)
    ids1, d1 = e.scan_topk(tables=t, k=k)
    e.set_option('scan_kernel', 0)
    rid, rd = O.scan_topk(t, codes, k)
    assert np.array_equal(ids, rid) and np.array_equal(bits(d), bits(rd))
    assert np.array_equal(ids1, rid) and np.array_equal(bits(d1), bits(rd))
    if N >= 32768 and B >= 64:                 # ... and it is what the dispatcher picks at this size
        ids0, d0 = e.scan_topk(queries=q, k=k)
        assert np.array_equal(ids0, rid) and np.array_equal(bits(d0), bits(rd))
