"""CPU: the C restatement (oracle/) reproduces what the compiled reference returned (tests/golden)."""
import numpy as np

import oracle as O


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_tables_bit_exact(golden):
    t = golden.query_tables_oracle()
    assert np.array_equal(bits(t[:8]), bits(golden.tables))


def test_single_query_table(golden):
    # pq_bind.precompute_adc_table is the L2 form regardless of metric (pq.py:200-224)
    Q = golden.Q
    if golden.metric == 'cosine':
        Q = O.l2_normalize(Q).astype(np.float32)
    t = O.adc_table(Q[1:2], golden.cb, 'euclidean')[0]
    assert np.array_equal(bits(t), bits(golden.table_single_q1))


def test_scan_bit_exact(golden):
    t = golden.query_tables_oracle()
    d = O.scan(t[0], golden.codes)
    assert np.array_equal(bits(d), bits(golden.scan_d))
    ids, dd = O.scan_topk(t[:4], golden.codes, 10)
    order = np.lexsort((np.arange(len(d)), d))[:10]
    assert np.array_equal(ids[0], order)
    assert np.array_equal(bits(dd[0]), bits(d[order]))


def test_knn_bit_exact(golden):
    g = golden.oracle_graph()
    t = golden.query_tables_oracle()
    l, d, found = O.hnsw_search(g, t, golden.k, golden.ef)
    assert (found == golden.k).all()
    assert np.array_equal(l, golden.knn_labels)
    assert np.array_equal(bits(d), bits(golden.knn_dists))


def test_filtered_bit_exact(golden):
    g = golden.oracle_graph()
    t = golden.query_tables_oracle()
    l, d, found = O.hnsw_search(g, t, golden.k, golden.ef, filter_labels=golden.allow)
    assert np.array_equal(l, golden.flt_labels)
    assert np.array_equal(bits(d), bits(golden.flt_dists))


def test_deleted_bit_exact(golden):
    g = golden.oracle_graph(deleted=True)
    t = golden.query_tables_oracle()
    l, d, found = O.hnsw_search(g, t, golden.k, golden.ef)
    assert np.array_equal(l, golden.del_labels)
    assert np.array_equal(bits(d), bits(golden.del_dists))
    assert not np.isin(l, golden.deleted).any()


def test_graph_views(golden):
    g = golden.oracle_graph()
    assert np.array_equal(g.codes(), golden.codes)
    assert np.array_equal(np.sort(g.labels()), np.sort(golden.labels))
    cnt, lk, dele = g.links0()
    assert cnt.max() <= g.max_M0 and not dele.any()
