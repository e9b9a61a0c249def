"""Hunt for a rare GPU-vs-oracle mismatch on host-built graphs (C5-like shape): loop builds for a time budget,
compare the fast walk with the oracle, and on any mismatch dump everything needed to replay it offline
(graph state, codebook, queries, both results) under gpurun_out/.  usage: repro_shapes.py SECONDS [threads]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import oracle as O  # noqa: E402
from annlite_b200.engine import Engine  # noqa: E402
from helpers import bits, tie_aware_rows  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = os.path.join(ROOT, 'gpurun_out')
os.makedirs(out, exist_ok=True)
N, D, M, Ks, k = 20000, 96, 16, 256, 10
rng = np.random.default_rng(5)
X = rng.standard_normal((N, D)).astype(np.float32)
Q0 = rng.standard_normal((96, D)).astype(np.float32)
ds = D // M
cb = np.stack([X[rng.choice(N, Ks, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
t_start, rep, nfail, rows = time.time(), 0, 0, 0
while time.time() - t_start < budget:
    e = Engine(D, M, Ks, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(N, M=16, ef_construction=100)
    tb = time.time()
    e.add_items(X, np.arange(N, dtype=np.uint64) + 11, num_threads=threads)
    tb = time.time() - tb
    st = e.get_graph()
    g = O.Graph.from_state(st, M, Ks)
    Q = np.concatenate([Q0, np.random.default_rng(1000 + rep).standard_normal((416, D)).astype(np.float32)])
    t = O.adc_table(Q, cb, 'euclidean')
    for ef in (256, 128, 200):
        l, d, s = e.search(tables=t, k=k, ef=ef, with_stats=True)
        ol, od, found, (hops, nbrs, evals) = O.hnsw_search(g, t, k, ef, with_counts=True)
        v = tie_aware_rows(l, d, ol, od)
        same = np.array([x == 'exact' for x in v])
        badrows = [i for i, x in enumerate(v) if x != 'exact'] + [int(i) for i in np.nonzero(same & ((s[:, 0] != hops) | (s[:, 1] != nbrs)))[0]]
        rows += len(v)
        if badrows:
            nfail += 1
            l2, d2, s2 = e.search(tables=t, k=k, ef=ef, with_stats=True)
            e.set_option('force_general', 2)
            l3, d3, s3 = e.search(tables=t, k=k, ef=ef, with_stats=True)
            e.set_option('force_general', 0)
            print('MISMATCH rep', rep, 'ef', ef, 'rows', badrows, 'verdicts', [v[i] for i in badrows],
                  'repeatable', bool(np.array_equal(l, l2) and np.array_equal(bits(d), bits(d2)) and np.array_equal(s, s2)),
                  'bitmap_walk_matches_oracle', tie_aware_rows(l3, d3, ol, od).count('diff') == 0,
                  'bitmap_rows', [tie_aware_rows(l3[i:i + 1], d3[i:i + 1], ol[i:i + 1], od[i:i + 1])[0] for i in badrows], flush=True)
            for i in badrows[:4]:
                print(' row', i, 'gpu', l[i].tolist(), d[i].tolist(), 'hops', s[i].tolist())
                print(' row', i, 'orc', ol[i].tolist(), od[i].tolist(), 'hops', int(hops[i]), int(nbrs[i]))
            np.savez_compressed(os.path.join(out, 'repro_%d_ef%d.npz' % (rep, ef)), cb=cb, Q=Q, ef=ef, k=k, rows=np.array(badrows),
                                gpu_l=l, gpu_d=d, gpu_s=s, gpu2_l=l2, gpu2_d=d2, bm_l=l3, bm_d=d3, orc_l=ol, orc_d=od, orc_hops=hops, orc_nbrs=nbrs,
                                **{'st_' + kk: np.asarray(vv) for kk, vv in st.items()})
    print('rep', rep, 'build %.2fs' % tb, 'maxlevel', g.maxlevel, 'fails so far', nfail, 'rows', rows, flush=True)
    rep += 1
    del e
print('DONE reps', rep, 'fails', nfail, 'rows', rows)
