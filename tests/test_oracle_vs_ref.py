"""CPU: pin the C restatement against the COMPILED REFERENCE on fresh seeded data (runs wherever
oracle/_ref exists -- it is built from /root/reference by oracle/build_ref.py and travels to the GPU
box as a binary).  Also checks the documented equivalence used to work around the reference's
knn_query segfault."""
import numpy as np
import pytest

import oracle as O
from oracle import ref_driver as R

pytestmark = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built')


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize('N,D,M,Ks,metric,seed,ties', [
    (3000, 32, 4, 256, 'euclidean', 21, False),
    (3000, 48, 8, 256, 'cosine', 22, False),
    (2000, 16, 4, 12, 'euclidean', 23, True),
    (2000, 24, 4, 400, 'inner_product', 24, False),
])
def test_oracle_equals_compiled_reference(N, D, M, Ks, metric, seed, ties):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((80, D)).astype(np.float32)
    if ties:
        X, Q = np.round(X), np.round(Q)
    ds = D // M
    cb = np.stack([X[rng.choice(N, Ks, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    codec = R.RefCodec(cb, metric)
    labels = rng.permutation(N).astype(np.uint64) * 3 + 1
    idx = R.RefHnswIndex(codec, metric, capacity=N, ef_search=48)
    idx.add_with_ids(X, labels, num_threads=1)
    Qp = idx._pre(Q)
    T = codec.get_dist_mat(Qp)
    To = O.adc_table(Qp, cb, metric)
    assert np.array_equal(bits(T), bits(To))
    g = O.Graph.from_state(idx.state(), M, Ks)
    for k, ef in [(10, 48), (1, 10), (30, 20)]:
        idx.ef_search = ef
        rl, rd = idx.knn_query(Q, k, num_threads=1, tables=T)
        ol, od, found = O.hnsw_search(g, To, k, ef)
        assert np.array_equal(rl, ol) and np.array_equal(bits(rd), bits(od))
        # multi-threaded and "filter = all ids" routes return the same thing
        rl8, rd8 = idx.knn_query(Q, k, num_threads=8, tables=T)
        assert np.array_equal(rl, rl8) and np.array_equal(bits(rd), bits(rd8))
        rlf, rdf = idx.knn_query(Q, k, indices=labels, tables=T)
        assert np.array_equal(rl, rlf) and np.array_equal(bits(rd), bits(rdf))
    idx.ef_search = 48
    allow = np.sort(labels[rng.random(N) < 0.4])
    rl, rd = idx.knn_query(Q, 10, indices=allow, tables=T)
    ol, od, _ = O.hnsw_search(g, To, 10, 48, filter_labels=allow)
    if np.isin(rl, allow).all():            # no binary-fuse false positive leaked into the reference result
        assert np.array_equal(rl, ol) and np.array_equal(bits(rd), bits(od))
    for l in labels[::11]:
        idx._index.mark_deleted(int(l))
    g = O.Graph.from_state(idx.state(), M, Ks)
    rl, rd = idx.knn_query(Q, 10, num_threads=1, tables=T)
    ol, od, _ = O.hnsw_search(g, To, 10, 48)
    assert np.array_equal(rl, ol) and np.array_equal(bits(rd), bits(od))
    # exhaustive scan + single-query table
    codes = g.codes()
    dref = np.asarray(R.pq_bind().dist_pqcodes_to_codebooks(T[3], codes), dtype=np.float32)
    assert np.array_equal(bits(dref), bits(O.scan(To[3], codes)))
    assert np.array_equal(bits(codec.precompute_adc(Qp[5])), bits(O.adc_table(Qp[5:6], cb, 'euclidean')[0]))


def test_ref_linear_scan_equals_oracle_topk():
    rng = np.random.default_rng(9)
    N, D, M, Ks = 4000, 32, 8, 256
    X = rng.standard_normal((N, D)).astype(np.float32)
    cb = np.stack([X[rng.choice(N, Ks, replace=False), m * 4:(m + 1) * 4] for m in range(M)]).astype(np.float32)
    codec = R.RefCodec(cb, 'euclidean')
    codes = codec.encode(X)
    for q in rng.standard_normal((5, D)).astype(np.float32):
        d, i = R.ref_pq_linear_scan(codec, codes, q, 10)
        oi, od = O.scan_topk(O.adc_table(q[None], cb), codes, 10)
        assert np.array_equal(bits(d.astype(np.float32)), bits(od[0]))
        assert np.array_equal(np.sort(i), np.sort(oi[0])) or len(set(od[0].tolist())) < 10
