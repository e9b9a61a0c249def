#!/usr/bin/env python
"""Full-size parity for BASELINE.json configs[1] (1M x 128d, M=8, ef=64, k=10), run on the GPU box:

1. the COMPILED REFERENCE (oracle/_ref) builds the index (all host threads) and answers 10 000 queries
   through knn_query;
2. its graph (Index.__getstate__) is adopted by the CUDA engine (annb_set_graph) and the same queries are
   answered by K1+K3;
3. ids / fp32 distances are compared row by row (tie-aware), and recall@10 of both against the exhaustive
   ADC ground truth (K2) and true L2 is reported.
Also compares recall of the index built by the product's own builder.  Writes one JSON to stdout.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from helpers import recall, tie_aware_rows  # noqa: E402
from oracle import ref_driver as R  # noqa: E402
from annlite_b200.engine import Engine  # noqa: E402


def main():
    sys.argv = [sys.argv[0]] + sys.argv[1:]
    a = bench.parse()
    ncores = os.cpu_count()
    cb = bench.train_codebook(a, bench.make_base(a, 0, 10_000))
    X = bench.make_base(a)
    Q = bench.make_queries(a, 1)[0]
    codec = R.RefCodec(cb, a.metric)
    idx = R.RefHnswIndex(codec, a.metric, capacity=a.n, ef_construction=a.efc, ef_search=a.ef, max_connection=a.M)
    t0 = time.time()
    idx.add_with_ids(X, np.arange(a.n), num_threads=ncores, batch=5000)
    t_ref_build = time.time() - t0
    tables = codec.get_dist_mat(idx._pre(Q))
    t0 = time.time()
    rl, rd = idx.knn_query(Q, a.k, num_threads=ncores, tables=tables)
    t_ref = time.time() - t0

    e = Engine(a.dim, a.m, a.ks, a.metric)
    e.set_codebook(cb)
    e.set_graph(idx.state())
    norm = 2 if a.metric == 'cosine' else 0
    gl, gd, st = e.search(queries=Q, k=a.k, ef=a.ef, normalize=norm, with_stats=True)
    v = tie_aware_rows(gl, gd, rl, rd)
    # table parity at full batch
    gt = e.adc_table(Q[:256], normalize=norm)
    tab_equal = bool(np.array_equal(gt.view(np.uint32), tables[:256].view(np.uint32)))
    # ground truths
    g = e.get_graph()
    n = g['cur_element_count']
    rec = g['data_level0'].reshape(n, -1)
    codes = np.ascontiguousarray(rec[:, g['offset_data']:g['label_offset']]).view(np.uint8).reshape(n, a.m)
    lab = np.ascontiguousarray(rec[:, g['label_offset']:g['label_offset'] + 8]).view(np.uint64).ravel()
    e.set_codes(codes)
    S = 2000
    gi, _ = e.scan_topk(tables=tables[:S], k=a.k)
    truth_adc = lab[gi]
    import torch
    Xd = torch.from_numpy(X).cuda()
    qd = torch.from_numpy(Q[:S]).cuda()
    d2 = (qd * qd).sum(1, keepdim=True) - 2 * qd @ Xd.T + (Xd * Xd).sum(1)[None]
    truth_l2 = d2.topk(a.k, dim=1, largest=False).indices.cpu().numpy().astype(np.uint64)
    out = {
        'config': f'{a.n}x{a.dim} {a.dist} {a.metric} M={a.m} Ks={a.ks} HNSW M={a.M} efc={a.efc} ef={a.ef} k={a.k}, {len(Q)} queries',
        'host_cores': ncores, 'ref_build_s': round(t_ref_build, 1), 'ref_knn_query_s': round(t_ref, 3),
        'tables_bit_equal_256': tab_equal,
        'rows': len(v), 'rows_exact': v.count('exact'), 'rows_tie_only': v.count('tie'), 'rows_diff': v.count('diff'),
        'max_rel_dist_err_on_equal_ids': float(np.max(np.abs(gd - rd)[gl == rl] / np.maximum(np.abs(rd[gl == rl]), 1e-30))) if (gl == rl).any() else None,
        'recall_ref_vs_adc': recall(rl[:S], truth_adc), 'recall_gpu_vs_adc': recall(gl[:S], truth_adc),
        'recall_ref_vs_l2': recall(rl[:S], truth_l2), 'recall_gpu_vs_l2': recall(gl[:S], truth_l2),
        'recall_gpu_vs_ref_ids': recall(gl, rl),
        'hops_per_query': float(st[:, 0].mean()), 'nbrs_per_query': float(st[:, 1].mean()),
    }
    # the product's own builder on the same data (thread-order dependent graph => compare recall only)
    e2 = Engine(a.dim, a.m, a.ks, a.metric)
    e2.set_codebook(cb)
    e2.init_graph(a.n, M=a.M, ef_construction=a.efc)
    t0 = time.time()
    e2.add_items(X if a.metric != 'cosine' else R.l2_normalize(X).astype(np.float32), np.arange(a.n, dtype=np.uint64), num_threads=ncores)
    out['our_build_s'] = round(time.time() - t0, 1)
    ol, od = e2.search(queries=Q, k=a.k, ef=a.ef, normalize=norm)
    out['recall_ourgraph_vs_adc'] = recall(ol[:S], truth_adc)
    out['recall_ourgraph_vs_l2'] = recall(ol[:S], truth_l2)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
