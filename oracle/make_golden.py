#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref) -- run in the build
container, where /root/reference exists; the fixtures are committed so they travel.

Each fixture holds seeded inputs (vectors, codebook, queries) and what the reference itself
returned for them through its own native entry points:

* ``tables_*``   pq_bind.batch_precompute_adc_table[_ip] (+ the `1/Ks - .` epilogue of
                 annlite/core/codec/pq.py:316-322), produced via RefCodec.get_dist_mat
* ``scan_d``     pq_bind.dist_pqcodes_to_codebooks for query 0 (all N codes)
* ``graph_*``    Index.__getstate__() after a single-threaded add_items (deterministic build)
* ``knn_*``      Index.knn_query, ``flt_*`` Index.knn_query_with_filter, ``del_*`` knn_query after
                 mark_deleted -- labels uint64 and fp32 distances exactly as returned.

Usage: python oracle/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_driver as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    # name, N, D, M, Ks, metric, seed, ties, ef, k, nq
    ('l2_m4', 2000, 32, 4, 256, 'euclidean', 11, False, 64, 10, 64),
    ('cos_m8', 2500, 64, 8, 256, 'cosine', 12, False, 64, 10, 64),
    ('ties_k16', 1500, 16, 4, 16, 'euclidean', 13, True, 32, 5, 64),
    ('ip_u16', 1500, 24, 4, 300, 'inner_product', 14, False, 50, 10, 48),
]


def kmeans_codebook(X, M, Ks, rng, iters=8):
    """Small deterministic Lloyd k-means (training is out of scope; any codebook is a valid input)."""
    N, D = X.shape
    ds = D // M
    cb = np.empty((M, Ks, ds), dtype=np.float32)
    for m in range(M):
        S = X[:, m * ds:(m + 1) * ds]
        C_ = S[rng.choice(N, Ks, replace=False)].copy()
        for _ in range(iters):
            d = ((S[:, None, :] - C_[None, :, :]) ** 2).sum(-1)
            a = d.argmin(1)
            for c in range(Ks):
                sel = S[a == c]
                if len(sel):
                    C_[c] = sel.mean(0)
        cb[m] = C_
    return cb


def make(name, N, D, M, Ks, metric, seed, ties, ef, k, nq):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((nq, D)).astype(np.float32)
    if ties:
        X, Q = np.round(X), np.round(Q)
    Xn = R.l2_normalize(X) if metric == 'cosine' else X
    cb = kmeans_codebook(Xn, M, Ks, rng)
    codec = R.RefCodec(cb, metric)
    labels = (rng.permutation(N).astype(np.uint64) * 2 + 5)  # non-trivial external labels
    idx = R.RefHnswIndex(codec, metric, capacity=N, ef_search=ef)
    idx.add_with_ids(X, labels, num_threads=1)
    st = idx.state()

    Qp = idx._pre(Q)                      # HnswIndex.pre_process: normalise (cosine)
    tables = codec.get_dist_mat(Qp)       # ... then get_dist_mat normalises again (pq.py:309)
    knn_l, knn_d = idx.knn_query(Q, k, num_threads=1, tables=tables)
    allow = np.sort(labels[rng.random(N) < 0.5])
    flt_l, flt_d = idx.knn_query(Q, k, indices=allow, tables=tables)
    assert np.isin(flt_l, allow).all(), 'binary-fuse false positive leaked; change the seed'
    codes = np.asarray(st['data_level0']).view(np.uint8).reshape(N, -1)[:, st['offset_data']:st['label_offset']]
    codes = np.ascontiguousarray(codes).view(codec.code_dtype).reshape(N, M)
    scan_d = np.asarray(R.pq_bind().dist_pqcodes_to_codebooks(tables[0], codes), dtype=np.float32)
    # single-query L2 table entry point (pq_bindings.pyx:85-145)
    single = codec.precompute_adc(Qp[1]).astype(np.float32)

    deleted = labels[::9].copy()
    for l in deleted:
        idx._index.mark_deleted(int(l))
    del_l, del_d = idx.knn_query(Q, k, num_threads=1, tables=tables)

    np.savez_compressed(
        os.path.join(OUT, name + '.npz'),
        X=X, Q=Q, codebook=cb, labels=labels, metric=np.array(metric), ef=ef, k=k,
        tables=tables[:8], table_single_q1=single, scan_d=scan_d, codes=codes,
        graph_level0=np.asarray(st['data_level0']).view(np.uint8),
        graph_links=np.asarray(st['link_lists']).view(np.uint8),
        graph_levels=np.asarray(st['element_levels']).astype(np.int32),
        graph_meta=np.array([st['size_data_per_element'], st['offset_data'], st['label_offset'],
                             st['size_links_per_element'], st['cur_element_count'], st['max_level'],
                             st['enterpoint_node'], st['max_M'], st['max_M0'], st['ef_construction'],
                             st['max_elements']], dtype=np.int64),
        graph_mult=np.float64(st['mult']),
        knn_labels=knn_l, knn_dists=knn_d, allow=allow, flt_labels=flt_l, flt_dists=flt_d,
        deleted=deleted, del_labels=del_l, del_dists=del_d)
    print(name, 'ok', os.path.getsize(os.path.join(OUT, name + '.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    for c in CASES:
        make(*c)
