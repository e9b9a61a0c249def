#!/usr/bin/env python
"""Side measurements for BASELINE.json configs[0] (C1: 10k exhaustive ADC scan) and configs[3]
(C4: 1M cosine + 50 % filter bitmap), GPU vs the compiled reference on the same box.  Not the headline
(bench.py is); results are committed under profiles/."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
import oracle as O  # noqa: E402
from helpers import recall, tie_aware_rows  # noqa: E402
from oracle import ref_driver as R  # noqa: E402
from annlite_b200.engine import Engine  # noqa: E402


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def c1():
    rng = np.random.default_rng(1)
    N, D, M, Ks, B, k = 10_000, 128, 8, 256, 1000, 10
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    from sklearn.cluster import KMeans
    cb = np.stack([KMeans(n_clusters=Ks, max_iter=20, n_init=1, random_state=0).fit(X[:, m * 16:(m + 1) * 16]).cluster_centers_
                   for m in range(M)]).astype(np.float32)
    codec = R.RefCodec(cb, 'euclidean')
    codes = codec.encode(X)
    e = Engine(D, M, Ks, 'euclidean')
    e.set_codebook(cb)
    e.set_codes(codes)
    ids, d = e.scan_topk(queries=Q, k=k)
    oi, od = O.scan_topk(O.adc_table(Q, cb), codes, k)
    t_gpu = timeit(lambda: e.scan_topk(queries=Q, k=k))
    import torch
    Qd = torch.from_numpy(Q).cuda()
    oi_d = torch.empty((B, k), dtype=torch.int64, device='cuda')
    od_d = torch.empty((B, k), dtype=torch.float32, device='cuda')
    t_dev = timeit(lambda: (e.scan_topk(queries=Qd, k=k, out_ids=oi_d, out_dists=od_d), e.sync()))
    kms = e.last_kernel_ms()
    # reference: PQIndex.search one query per call (annlite/core/index/pq_index.py:29-56), bounded sample
    S = 200
    t0 = time.perf_counter()
    ref_ids = [R.ref_pq_linear_scan(codec, codes, Q[i], k)[1] for i in range(S)]
    t_ref = time.perf_counter() - t0
    ref_match = float(np.mean([set(a.tolist()) == set(b.tolist()) for a, b in zip(ref_ids, ids[:S])]))
    return {'config': 'C1 10k x 128d, M=8, exhaustive ADC, k=10, 1000 queries',
            'ids_exact_vs_oracle': bool(np.array_equal(ids, oi)), 'dists_bit_exact': bool(np.array_equal(d.view(np.uint32), od.view(np.uint32))),
            'id_sets_equal_vs_reference_sample': ref_match,
            'gpu_qps_host_buffers': B / t_gpu, 'gpu_qps_device_buffers': B / t_dev, 'scan_kernel_ms': kms['scan_ms'],
            'table_kernel_ms': kms['table_ms'], 'reference_qps_one_query_per_call': S / t_ref,
            'alg_bytes_per_query': N * M}


def c4(n=1_000_000, correlated=False):
    """correlated=False: the filtering_bench.py shape, a random 50 % bitmap.  correlated=True: three gaussian blobs
    (examples/pq_benchmark.py:25-28 data) and a filter that admits two WHOLE blobs (s ~ 0.67): a third of the queries
    start inside the excluded blob and must walk out of it before they find any admissible node -- the case where
    the flagged walk's list overflows and those queries are re-run on the bitmap walk."""
    a = bench.parse(['--metric', 'cosine', '--n', str(n)] + (['--dist', 'blobs'] if correlated else []))
    ncores = os.cpu_count()
    cb = bench.train_codebook(a, bench.make_base(a, 0, 10_000))
    X = bench.make_base(a)
    Q = bench.make_queries(a, 1)[0]
    rng = np.random.default_rng(4)
    if correlated:
        blob = np.random.default_rng([2, 77]).integers(0, 3, a.n)      # the labels bench.make_base draws
        allow = np.nonzero(blob != 2)[0].astype(np.uint64)
    else:
        allow = np.nonzero(rng.random(a.n) < 0.5)[0].astype(np.uint64)
    e = Engine(a.dim, a.m, a.ks, 'cosine')
    e.set_codebook(cb)
    e.init_graph(a.n, M=a.M, ef_construction=a.efc)
    t0 = time.time()
    e.add_items(R.l2_normalize(X).astype(np.float32), np.arange(a.n, dtype=np.uint64), num_threads=ncores)
    t_build = time.time() - t0
    l, d, st = e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, filter_labels=allow, with_stats=True)
    t_flt = timeit(lambda: e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, filter_labels=allow), 3)
    e.set_option('chunks', 1)                                      # one launch over the whole batch: the kernel's own time
    e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, filter_labels=allow)
    k_ms = e.last_kernel_ms()['search_ms']
    e.set_option('chunks', 0)
    t_plain = timeit(lambda: e.search(queries=Q, k=a.k, ef=a.ef, normalize=2), 3)
    fb_batches, fb_queries = e.fallback_count, e.fallback_queries
    e.set_option('chunks', 1)                                      # one launch: the plain walk's kernel time on the same index
    _, _, st_plain = e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, with_stats=True)
    e.search(queries=Q, k=a.k, ef=a.ef, normalize=2)
    k_ms_plain = e.last_kernel_ms()['search_ms']
    e.set_option('chunks', 0)
    # A/B: round 1's flagged walk (one list with PASS flags, shared-memory merge) on the same call
    e.set_option('flagged_kernel', 1)
    l1, d1 = e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, filter_labels=allow)
    t_flt1 = timeit(lambda: e.search(queries=Q, k=a.k, ef=a.ef, normalize=2, filter_labels=allow), 3)
    k_ms1 = e.last_kernel_ms()['search_ms']                        # (round 1's kernel always runs as one launch)
    e.set_option('flagged_kernel', 0)
    same_as_round1 = tie_aware_rows(l, d, l1, d1).count('diff')
    # streamed form (annb_search_submit_filtered): pinned host buffers, two batches in flight, the filter label list
    # uploaded and turned into the by-id bitmap per batch on the batch's own lane
    import torch
    Qp = torch.from_numpy(Q).pin_memory().numpy()
    fl = torch.from_numpy(allow.view(np.int64)).pin_memory().numpy().view(np.uint64)
    outs = [(torch.empty((len(Q), a.k), dtype=torch.int64).pin_memory().numpy().view(np.uint64),
             torch.empty((len(Q), a.k), dtype=torch.float32).pin_memory().numpy()) for _ in range(2)]

    def streamed(nb):
        tk = []
        t0 = time.perf_counter()
        for i in range(nb):
            if i >= 2:
                e.search_wait(tk[i - 2])
            tk.append(e.search_submit(Qp, outs[i & 1][0], outs[i & 1][1], k=a.k, ef=a.ef, normalize=2, filter_labels=fl))
        e.search_wait(tk[-2])
        e.search_wait(tk[-1])
        return (time.perf_counter() - t0) / nb
    streamed(4)
    t_str = min(streamed(20) for _ in range(3)) if not correlated else streamed(4)
    streamed_same = bool(np.array_equal(outs[1][0], l) and np.array_equal(outs[1][1], d))
    # parity on the same graph: oracle port (filter semantic), bounded sample
    S = 2000
    g = O.Graph.from_state(e.get_graph(), a.m, a.ks)
    tq = O.adc_table(O.l2_normalize(Q[:S]).astype(np.float32), cb, 'cosine')
    ol, od, found = O.hnsw_search(g, tq, a.k, a.ef, filter_labels=allow)
    verdict = tie_aware_rows(l[:S], d[:S], ol, od)
    rel = np.abs(d[:S] - od)[l[:S] == ol] / np.maximum(np.abs(od[l[:S] == ol]), 1e-12)
    out = {'config': f'C4 {a.n} x 128d cosine {a.dist}, {"two whole blobs admitted" if correlated else "random 50% filter bitmap"} '
                     f'({len(allow)} ids), M=8, HNSW ef=64 k=10, {len(Q)} queries',
           'index_build_s': t_build, 'gpu_filtered_qps_host_buffers': len(Q) / t_flt, 'gpu_filtered_kernel_ms': k_ms,
           'flagged_walk_fallback_batches': fb_batches, 'flagged_walk_fallback_queries': fb_queries,
           'searches_run': 5,
           'filtered_kernel': 'hnsw_walk4f (two register lists, table built in shared memory)',
           'gpu_filtered_qps_streamed_host_buffers': len(Q) / t_str, 'streamed_ms_per_batch': t_str * 1e3,
           'streamed_rows_equal_blocking_call': streamed_same,
           'round1_flagged_walk': {'gpu_filtered_qps_host_buffers': len(Q) / t_flt1, 'kernel_ms_excl_table_kernel': k_ms1,
                                   'rows_diff_vs_new_kernel': same_as_round1},
           'gpu_unfiltered_qps_host_buffers': len(Q) / t_plain, 'unfiltered_kernel_ms': k_ms_plain,
           'unfiltered_hops_per_query': float(st_plain[:, 0].mean()), 'all_results_pass_filter': bool(np.isin(l, allow).all()),
           'hops_per_query': float(st[:, 0].mean()), 'evals_per_query': float(st[:, 2].mean()),
           'parity_sample': S, 'rows_exact': verdict.count('exact'), 'rows_tie': verdict.count('tie'),
           'rows_diff': verdict.count('diff'), 'recall_vs_oracle_ids': recall(l[:S], ol),
           'max_rel_dist_err_equal_ids': float(rel.max()) if rel.size else None}
    # reference: knn_query_with_filter on its own (thread-order dependent) graph, bounded sample
    if R.available() and n <= 1_000_000 and not correlated and not os.environ.get('C4_SKIP_REF'):
        codec = R.RefCodec(cb, 'cosine')
        idx = R.RefHnswIndex(codec, 'cosine', capacity=a.n, ef_construction=a.efc, ef_search=a.ef, max_connection=a.M)
        t0 = time.time()
        idx.add_with_ids(X, np.arange(a.n), num_threads=ncores, batch=5000)
        out['ref_build_s'] = time.time() - t0
        Sr = 2000
        t0 = time.perf_counter()
        rl, rd = idx.knn_query(Q[:Sr], a.k, num_threads=ncores, indices=allow)
        out['reference_filtered_qps_batched'] = Sr / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for j in range(5):
            idx.search(Q[j], limit=a.k, indices=allow)
        out['reference_filtered_qps_one_per_call'] = 5 / (time.perf_counter() - t0)
    return out


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'c1'
    sys.argv = sys.argv[:1]
    print(json.dumps(c1() if which == 'c1' else c4(int(os.environ.get('C4_N', 1_000_000)), correlated=(which == 'c4corr'))))
