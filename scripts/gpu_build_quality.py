#!/usr/bin/env python
"""Build time and graph quality (recall@10 vs the exhaustive ADC scan) of the GPU builder against the host builder
on the same rows, for several batch fractions.  One JSON line per build."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from annlite_b200.engine import Engine
from helpers import recall


def blobs(n, d, seed, centers=64):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centers, d)).astype(np.float32) * 3
    return (c[rng.integers(0, centers, n)] + rng.standard_normal((n, d)).astype(np.float32)).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--base-n', dest='n', type=int, default=200_000)
    ap.add_argument('--dim', type=int, default=64)
    ap.add_argument('--pq-m', dest='m', type=int, default=8)
    ap.add_argument('--dist', default='blobs')
    ap.add_argument('--fracs', default='4,8,16,32')
    ap.add_argument('--efc', type=int, default=200)
    ap.add_argument('--host', type=int, default=1)
    a = ap.parse_args()
    X = blobs(a.n, a.dim, 5) if a.dist == 'blobs' else np.random.default_rng(5).standard_normal((a.n, a.dim), dtype=np.float32)
    Q = blobs(1000, a.dim, 6) if a.dist == 'blobs' else np.random.default_rng(6).standard_normal((1000, a.dim), dtype=np.float32)
    ds = a.dim // a.m
    from sklearn.cluster import KMeans
    cb = np.stack([KMeans(256, n_init=1, max_iter=10, random_state=0).fit(X[:20000, m * ds:(m + 1) * ds]).cluster_centers_ for m in range(a.m)]).astype(np.float32)
    truth = None
    variants = ([('host', 0)] if a.host else []) + [('gpu', int(f)) for f in a.fracs.split(',')]
    for kind, frac in variants:
        e = Engine(a.dim, a.m, 256, 'euclidean')
        e.set_codebook(cb)
        e.set_option('gpu_build', 1 if kind == 'gpu' else 0)
        if frac:
            e.set_option('gpu_build_frac', frac)
        e.init_graph(a.n, M=16, ef_construction=a.efc)
        t0 = time.time()
        e.add_items(X, np.arange(a.n, dtype=np.uint64))
        tb = time.time() - t0
        if truth is None:
            g = e.get_graph()
            rec = g['data_level0'].reshape(a.n, -1)
            codes = np.ascontiguousarray(rec[:, g['offset_data']:g['label_offset']]).view(np.uint8).reshape(a.n, a.m)
            lab = np.ascontiguousarray(rec[:, g['label_offset']:g['label_offset'] + 8]).view(np.uint64).ravel()
            order = np.argsort(lab)
            e.set_codes(codes[order])          # row i = label i
            gi, _ = e.scan_topk(queries=Q, k=10)
            truth = gi.astype(np.uint64)
        out = {'builder': kind, 'frac': frac, 'n': a.n, 'dim': a.dim, 'M': a.m, 'build_s': round(tb, 2)}
        for ef in (32, 64, 128):
            l, d, st = e.search(queries=Q, k=10, ef=ef, with_stats=True)
            out[f'recall_ef{ef}'] = round(recall(l, truth), 4)
            out[f'hops_ef{ef}'] = round(float(st[:, 0].mean()), 1)
        g = e.get_graph()
        cnt = np.ascontiguousarray(g['data_level0'].reshape(a.n, -1)[:, 0:2]).view(np.uint16).ravel()
        out['mean_degree'] = round(float(cnt.mean()), 2)
        print(json.dumps(out), flush=True)
        e.close()


if __name__ == '__main__':
    main()
