#!/usr/bin/env python
"""QPS-vs-recall sweeps for the BASELINE.json shapes that do not fit a round at full size, scaled in N:
  C5-like: N x 96d, M=16 (ds=6), batch 4096, ef in {16,32,64,128,256}, k=10
  C3-like: N x 768d, M=32 (ds=24), ef=128, k=100
One GPU, index built by the product's builder; QPS through the streamed host-buffer API
(annb_search_submit/wait, H2D + D2H inside the timed region) and device-resident; recall@k against the
exhaustive ADC scan (K2) and against true L2.  Writes one JSON object to stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench  # noqa: E402
from helpers import recall  # noqa: E402
from annlite_b200.engine import Engine  # noqa: E402


def run(n, dim, m, batch, efs, k, steps=30):
    import torch
    sys.argv = ['x', '--n', str(n), '--dim', str(dim), '--m', str(m), '--batch', str(batch), '--k', str(k)]
    a = bench.parse()
    cb = bench.train_codebook(a, bench.make_base(a, 0, 10_000))
    X = bench.make_base(a)
    Q = bench.make_queries(a, 2)
    e = Engine(dim, m, a.ks, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(n, M=a.M, ef_construction=a.efc)
    t0 = time.time()
    e.add_items(X, np.arange(n, dtype=np.uint64))
    build_s = time.time() - t0
    g = e.get_graph()
    rec = g['data_level0'].reshape(n, -1)
    codes = np.ascontiguousarray(rec[:, g['offset_data']:g['label_offset']]).view(np.uint8).reshape(n, m)
    e.set_codes(codes)
    S = min(1000, batch)
    gi, _ = e.scan_topk(queries=Q[0][:S], k=k)
    Xd = torch.from_numpy(X).cuda()
    qd = torch.from_numpy(Q[0][:S]).cuda()
    d2 = (qd * qd).sum(1, keepdim=True) - 2 * qd @ Xd.T + (Xd * Xd).sum(1)[None]
    gt = d2.topk(k, dim=1, largest=False).indices.cpu().numpy().astype(np.uint64)
    del Xd, d2
    Qp = torch.from_numpy(Q).pin_memory()
    Qd = torch.from_numpy(Q).cuda()
    hl = [torch.empty((batch, k), dtype=torch.int64).pin_memory() for _ in range(2)]
    hd = [torch.empty((batch, k), dtype=torch.float32).pin_memory() for _ in range(2)]
    dl = [torch.empty((batch, k), dtype=torch.int64, device='cuda') for _ in range(2)]
    dd = [torch.empty((batch, k), dtype=torch.float32, device='cuda') for _ in range(2)]
    out = {'config': f'{n}x{dim} fp32 gaussian, M={m} Ks={a.ks}, HNSW M={a.M} efc={a.efc}, k={k}, batch={batch}',
           'index_build_s': round(build_s, 1), 'points': []}

    def streamed(host, ef):
        pend = []
        for i in range(steps + 4):
            if i == 4:
                while pend:
                    e.search_wait(pend.pop(0))
                e.sync()
                t0 = time.perf_counter()
            if len(pend) == 2:
                e.search_wait(pend.pop(0))
            if host:
                pend.append(e.search_submit(Qp.numpy()[i & 1], hl[i & 1].numpy().view(np.uint64), hd[i & 1].numpy(), k=k, ef=ef))
            else:
                pend.append(e.search_submit(Qd[i & 1], dl[i & 1], dd[i & 1], k=k, ef=ef))
        while pend:
            e.search_wait(pend.pop(0))
        e.sync()
        return batch * steps / (time.perf_counter() - t0)

    for ef in efs:
        l, d, st = e.search(queries=Q[0], k=k, ef=ef, with_stats=True)
        kms = e.last_kernel_ms()['search_ms']
        out['points'].append({'ef': ef, 'qps_e2e_streamed': round(streamed(True, ef), 1),
                              'qps_resident_streamed': round(streamed(False, ef), 1), 'blocking_call_search_span_ms': round(kms, 4),
                              'recall_vs_exhaustive_adc': round(recall(l[:S], gi.astype(np.uint64)), 4),
                              'recall_vs_true_l2': round(recall(l[:S], gt), 4),
                              'hops_per_query': round(float(st[:, 0].mean()), 1),
                              'nbrs_per_query': round(float(st[:, 1].mean()), 1)})
    return out


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'c5'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    if which == 'c5':
        res = run(n, 96, 16, 4096, [16, 32, 64, 128, 256], 10)
    else:
        res = run(n, 768, 32, 4096, [128], 100)
    print(json.dumps(res))
