"""GPU: size-independent properties of the search path on an index built by the product itself
(200k x 64d, M=8 -- large enough that the walk records exceed L2 slices and queries take ~100 hops),
where the oracle would be too slow to be the checker for every query."""
import numpy as np
import pytest

import oracle as O
from annlite_b200.engine import Engine
from helpers import bits, recall

pytestmark = pytest.mark.gpu

N, D, M, KS, B = 200_000, 64, 8, 256, 4096


@pytest.fixture(scope='module')
def big():
    rng = np.random.default_rng(42)
    # i.i.d. gaussian: continuous values, so exact fp32 distance ties (the only freedom the GPU walk
    # takes relative to the reference) are practically absent and the checks below can be strict
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((B, D)).astype(np.float32)
    ds = D // M
    cb = np.stack([X[rng.choice(N, KS, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, KS, 'euclidean')
    e.set_codebook(cb)
    e.init_graph(N, M=16, ef_construction=100)
    labels = (np.arange(N, dtype=np.uint64) * 7 + 3)
    e.add_items(X, labels)
    g = e.get_graph()
    rec = g['data_level0'].reshape(N, -1)
    codes = np.ascontiguousarray(rec[:, g['offset_data']:g['label_offset']])
    lab = np.ascontiguousarray(rec[:, g['label_offset']:g['label_offset'] + 8]).view(np.uint64).ravel()
    return dict(e=e, X=X, Q=Q, cb=cb, labels=labels, codes=codes, lab_by_id=lab)


def test_rows_sorted_unique_and_distances_are_exact_adc(big):
    e = big['e']
    l, d, st = e.search(queries=big['Q'], k=10, ef=64, with_stats=True)
    assert (np.diff(d, axis=1) >= 0).all()
    assert all(len(set(r.tolist())) == 10 for r in l)
    assert np.isin(l, big['labels']).all()
    # checksum: every returned distance is the sequential fp32 ADC sum of that label's stored code
    T = e.adc_table(big['Q'])
    inv = {int(v): i for i, v in enumerate(big['lab_by_id'])}
    for b in range(0, B, 97):
        for j in range(10):
            code = big['codes'][inv[int(l[b, j])]]
            acc = np.float32(0)
            for m in range(M):
                acc = np.float32(acc + T[b, m, code[m]])
            assert acc.view(np.uint32) == d[b, j].view(np.uint32)
    assert (st[:, 0] >= 1).all() and (st[:, 1] <= st[:, 0] * 32).all() and (st[:, 1] >= 10).all()


def test_idempotent_and_buffer_spaces_agree(big):
    torch = pytest.importorskip('torch')
    e = big['e']
    l1, d1 = e.search(queries=big['Q'], k=10, ef=64)
    l2, d2 = e.search(queries=big['Q'], k=10, ef=64)
    assert np.array_equal(l1, l2) and np.array_equal(bits(d1), bits(d2))
    qd = torch.from_numpy(big['Q']).cuda()
    ol = torch.empty((B, 10), dtype=torch.int64, device='cuda')
    od = torch.empty((B, 10), dtype=torch.float32, device='cuda')
    e.search(queries=qd, k=10, ef=64, out_labels=ol, out_dists=od)
    e.sync()
    assert np.array_equal(ol.cpu().numpy().view(np.uint64), l1) and np.array_equal(bits(od.cpu().numpy()), bits(d1))
    # tables handed over explicitly (the reference's dtables argument) == fused path
    l3, d3 = e.search(tables=e.adc_table(big['Q']), k=10, ef=64)
    assert np.array_equal(l1, l3) and np.array_equal(bits(d1), bits(d3))


def test_general_walk_with_all_pass_filter_equals_fast_walk(big):
    e = big['e']
    l1, d1, s1 = e.search(queries=big['Q'][:1024], k=10, ef=64, with_stats=True)
    l2, d2, s2 = e.search(queries=big['Q'][:1024], k=10, ef=64, filter_labels=big['labels'], with_stats=True)
    same = (l1 == l2).all(1)
    assert same.mean() >= 0.999          # differences only through exact-distance ties
    assert np.array_equal(s1[same, 0], s2[same, 0]) and np.array_equal(s1[same, 1], s2[same, 1])
    assert (s2[:, 2] <= s1[:, 2]).all()  # exact visited set => never more distance evaluations


def test_filter_and_delete_semantics(big):
    e = big['e']
    allow = big['labels'][::3]
    l, d = e.search(queries=big['Q'][:512], k=10, ef=64, filter_labels=allow)
    assert np.isin(l, allow).all() and (np.diff(d, axis=1) >= 0).all()
    l0, _ = e.search(queries=big['Q'][:64], k=10, ef=64)
    victims = np.unique(l0[:, 0])
    for v in victims:
        e.mark_deleted(int(v))
    l1, _ = e.search(queries=big['Q'][:64], k=10, ef=64)
    assert not np.isin(l1, victims).any()
    for v in victims:
        e.unmark_deleted(int(v))
    l2, _ = e.search(queries=big['Q'][:64], k=10, ef=64)
    assert np.array_equal(l0, l2)


def test_recall_grows_with_ef_and_k_prefix_property(big):
    e = big['e']
    q = big['Q'][:512]
    e.set_codes(big['codes'])
    gi, _ = e.scan_topk(queries=q, k=10)
    truth = big['lab_by_id'][gi]
    r = [recall(e.search(queries=q, k=10, ef=ef)[0], truth) for ef in (10, 40, 160, 500)]
    assert r[0] <= r[1] + 0.01 <= r[2] + 0.02 <= r[3] + 0.03 and r[3] > r[0]
    # k results are the k nearest of the ef list: top-5 of (k=10) == (k=5) at the same ef
    l10, d10 = e.search(queries=q, k=10, ef=64)
    l5, d5 = e.search(queries=q, k=5, ef=64)
    assert np.array_equal(bits(d10[:, :5]), bits(d5))


def test_save_load_roundtrip_same_answers(big, tmp_path):
    e = big['e']
    p = tmp_path / 'big.hnsw'
    e.save_index(p)
    e2 = Engine(D, M, KS, 'euclidean')
    e2.set_codebook(big['cb'])
    e2.load_index(p)
    l1, d1 = e.search(queries=big['Q'][:1024], k=10, ef=64)
    l2, d2 = e2.search(queries=big['Q'][:1024], k=10, ef=64)
    assert np.array_equal(l1, l2) and np.array_equal(bits(d1), bits(d2))


def test_oracle_agrees_on_a_sample(big):
    e = big['e']
    g = O.Graph.from_state(e.get_graph(), M, KS)
    q = big['Q'][:200]
    t = O.adc_table(q, big['cb'])
    ol, od, _, (hops, nbrs, _), ties = O.hnsw_search(g, t, 10, 64, with_counts=True, with_ties=True)
    l, d, st = e.search(queries=q, k=10, ef=64, with_stats=True)
    same = (l == ol).all(1)
    assert same.mean() >= 0.99 and same[ties == 0].all()
    assert np.array_equal(bits(d[same]), bits(od[same]))
    same &= ties == 0                     # a tie the reference resolved by heap order can cost/save a hop
    assert np.array_equal(st[same, 0], hops[same]) and np.array_equal(st[same, 1], nbrs[same])


def test_streamed_submit_wait_matches_blocking_calls(big):
    torch = pytest.importorskip('torch')
    e = big['e']
    Q = big['Q']
    ref = [e.search(queries=Q[i * 512:(i + 1) * 512], k=10, ef=64) for i in range(6)]
    # host buffers, two batches in flight
    outs = [(np.empty((512, 10), np.uint64), np.empty((512, 10), np.float32)) for _ in range(6)]
    tickets = []
    for i in range(6):
        if len(tickets) == 2:
            e.search_wait(tickets.pop(0))
        tickets.append(e.search_submit(Q[i * 512:(i + 1) * 512], outs[i][0], outs[i][1], k=10, ef=64))
    while tickets:
        e.search_wait(tickets.pop(0))
    for (l, d), (rl, rd) in zip(outs, ref):
        assert np.array_equal(l, rl) and np.array_equal(bits(d), bits(rd))
    # device buffers
    qd = torch.from_numpy(Q).cuda()
    dl = [torch.empty((512, 10), dtype=torch.int64, device='cuda') for _ in range(6)]
    dd = [torch.empty((512, 10), dtype=torch.float32, device='cuda') for _ in range(6)]
    t = [e.search_submit(qd[i * 512:(i + 1) * 512], dl[i], dd[i], k=10, ef=64) for i in range(6)]   # a 3rd submit waits for its lane
    for x in t:
        e.search_wait(x)
    e.sync()
    for i, (rl, rd) in enumerate(ref):
        assert np.array_equal(dl[i].cpu().numpy().view(np.uint64), rl) and np.array_equal(bits(dd[i].cpu().numpy()), bits(rd))
    # a blocking call in between streamed ones is fine
    t0 = e.search_submit(Q[:512], outs[0][0], outs[0][1], k=10, ef=64)
    l, d = e.search(queries=Q[512:1024], k=10, ef=64)
    e.search_wait(t0)
    assert np.array_equal(l, ref[1][0]) and np.array_equal(outs[0][0], ref[0][0])
    # the forced filtered / deletion-aware route is served by the streamed form too (hnsw_walk4f): the same rows
    e.set_option('force_general', 1)
    try:
        t1 = e.search_submit(Q[:512], outs[1][0], outs[1][1], k=10, ef=64)
        e.search_wait(t1)
    finally:
        e.set_option('force_general', 0)
    assert np.array_equal(outs[1][0], ref[0][0]) and np.array_equal(bits(outs[1][1]), bits(ref[0][1]))
    # ... but not by the bitmap walk
    e.set_option('force_general', 2)
    try:
        with pytest.raises(RuntimeError, match='bitmap walk'):
            e.search_submit(Q[:8], outs[0][0][:8], outs[0][1][:8], k=10, ef=64)
    finally:
        e.set_option('force_general', 0)
