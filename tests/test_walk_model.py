"""CPU: executable statement of the two equivalence arguments the CUDA walks rest on.

A plain-Python model of (1) the single-list, visited-free walk (hnsw_walk_fast) and (2) the flagged
single-list walk for filters / deletions (hnsw_walk_flagged) is run against the oracle -- which is
bit-identical to the reference's two-heap + visited-array search -- on the committed fixtures.  The models
use no visited set and no candidate heap; they must return the same ids, distances AND the same number of
hops (every expansion the reference makes, the model makes)."""
import numpy as np
import pytest

import oracle as O

FLT_MAX = np.float32(3.4028234663852886e+38)


class G:
    def __init__(self, g):
        self.g = g
        self.cnt, self.lk, self.dele = g.links0()
        self.codes = g.codes()
        self.labels = g.labels()

    def dist(self, t, i):
        r = np.float32(0)
        for m in range(self.g.M_sub):
            r = np.float32(r + t[m, self.codes[i, m]])
        return r

    def descend(self, t):
        g = self.g
        cur, cd, hops = g.enterpoint, self.dist(t, g.enterpoint), 0
        for level in range(g.maxlevel, 0, -1):
            changed = True
            while changed:
                changed, hops = False, hops + 1
                o = int(g.link_off[cur]) + (level - 1) * g.size_links_per_elem
                c = int(np.frombuffer(g.links[o:o + 2].tobytes(), dtype=np.uint16)[0])
                for x in np.frombuffer(g.links[o + 4:o + 4 + 4 * c].tobytes(), dtype=np.uint32):
                    d = self.dist(t, int(x))
                    if d < cd:
                        cd, cur, changed = d, int(x), True
        return cur, cd, hops


def walk_single_list(G_, t, k, ef):
    """hnsw_walk_fast: ONE sorted list of <= ef (d, id, expanded); no visited set, no candidate heap."""
    cur, cd, hops = G_.descend(t)
    L = [(cd, cur, True)]
    node = cur
    while True:
        hops += 1
        worst = L[ef - 1][0] if len(L) >= ef else np.float32(np.inf)
        listed = {e[1] for e in L}
        for j in range(G_.cnt[node]):
            x = int(G_.lk[node, j])
            d = G_.dist(t, x)
            if d < worst and x not in listed:                 # admission (:306) + "still listed" check
                pos = sum(1 for e in L if e[0] <= d)           # after equal keys, arrival order
                L.insert(pos, (d, x, False))
                listed.add(x)
        L = L[:ef]                                             # the tail beyond ef falls off
        nxt = next((i for i, e in enumerate(L) if not e[2]), None)
        if nxt is None:
            break
        node = L[nxt][1]
        L[nxt] = (L[nxt][0], node, True)
    top = L[:k]
    return np.array([G_.labels[e[1]] for e in top], dtype=np.uint64), np.array([e[0] for e in top], dtype=np.float32), hops


def walk_flagged(G_, t, k, ef, passes, use_filter, has_del):
    """hnsw_walk_flagged: one list of every candidate with a PASS flag; lowerBound read off the list."""
    cur, cd, hops = G_.descend(t)
    p0 = passes(cur)
    L = [(cd, cur, True, p0)]
    lb = cd if p0 else FLT_MAX
    npass = 1 if p0 else 0
    node = cur
    while True:
        hops += 1
        listed = {e[1] for e in L}
        for j in range(G_.cnt[node]):
            x = int(G_.lk[node, j])
            d = G_.dist(t, x)
            if (npass < ef or d < lb) and x not in listed:     # :306 / :413 against the hop-start lowerBound
                pos = sum(1 for e in L if e[0] <= d)
                L.insert(pos, (d, x, False, passes(x)))
                listed.add(x)
        idx = [i for i, e in enumerate(L) if e[3]]
        if len(idx) >= ef:
            L = L[:idx[ef - 1] + 1]                            # everything behind the ef-th passing entry is dead
            npass, lb = ef, L[-1][0]
        else:
            npass = len(idx)
            if idx:
                lb = L[idx[-1]][0]
        nxt = next((i for i, e in enumerate(L) if not e[2]), None)
        if nxt is None:
            break
        key = L[nxt][0]
        if use_filter:
            if key > lb:
                break                                          # :371
        elif key > lb and (npass == ef or not has_del):
            break                                              # :270
        node = L[nxt][1]
        L[nxt] = (L[nxt][0], node, True, L[nxt][3])
    top = [e for e in L if e[3]][:k]
    return np.array([G_.labels[e[1]] for e in top], dtype=np.uint64), np.array([e[0] for e in top], dtype=np.float32), hops


NQ = 12


def _same(l, d, rl, rd):
    order = np.lexsort((l, d))                 # rows come back ascending (dist, label)
    return np.array_equal(l[order], rl) and np.array_equal(d[order].view(np.uint32), rd.view(np.uint32))


def test_single_list_walk_equals_two_heap_walk(golden):
    g = golden.oracle_graph()
    t = golden.query_tables_oracle()[:NQ]
    rl, rd, found, (hops, nbrs, evals) = O.hnsw_search(g, t, golden.k, golden.ef, with_counts=True)
    M = G(g)
    bad = 0
    for b in range(NQ):
        l, d, h = walk_single_list(M, t[b], golden.k, golden.ef)
        ok = _same(l, d, rl[b], rd[b]) and h == hops[b]
        bad += not ok
    assert bad <= (3 if golden.name == 'ties_k16' else 0)      # only exact fp32 ties may differ


def test_flagged_walk_equals_filtered_two_heap_walk(golden):
    g = golden.oracle_graph()
    t = golden.query_tables_oracle()[:NQ]
    rl, rd, found, (hops, _, _) = O.hnsw_search(g, t, golden.k, golden.ef, filter_labels=golden.allow, with_counts=True)
    M = G(g)
    allowed = set(golden.allow.tolist())
    bad = 0
    for b in range(NQ):
        l, d, h = walk_flagged(M, t[b], golden.k, golden.ef, lambda i: int(M.labels[i]) in allowed, True, False)
        bad += not (_same(l, d, rl[b], rd[b]) and h == hops[b])
    assert bad <= (3 if golden.name == 'ties_k16' else 0)


def test_flagged_walk_equals_deletion_aware_two_heap_walk(golden):
    g = golden.oracle_graph(deleted=True)
    t = golden.query_tables_oracle()[:NQ]
    rl, rd, found, (hops, _, _) = O.hnsw_search(g, t, golden.k, golden.ef, with_counts=True)
    M = G(g)
    bad = 0
    for b in range(NQ):
        l, d, h = walk_flagged(M, t[b], golden.k, golden.ef, lambda i: not M.dele[i], False, True)
        bad += not (_same(l, d, rl[b], rd[b]) and h == hops[b])
    assert bad <= (3 if golden.name == 'ties_k16' else 0)


# ---- the same equivalence at the sizes the GPU shape tests use, through the scalar C model ---------------
def _host_built(N, D, M, Ks, seed, threads, Mconn=16, efc=100):
    from annlite_b200.engine import Engine
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((N, D)).astype(np.float32)
    Q = rng.standard_normal((256, D)).astype(np.float32)
    ds = D // M
    cb = np.stack([X[rng.choice(N, Ks, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    e = Engine(D, M, Ks, 'euclidean', device=-1)          # host-only handle: graph building needs no GPU
    e.init_graph(N, M=Mconn, ef_construction=efc)
    e.add_items_with_tables(O.encode(X, cb), O.adc_table(X, cb), np.arange(N, dtype=np.uint64) + 11, num_threads=threads)
    return O.Graph.from_state(e.get_graph(), M, Ks), O.adc_table(Q, cb)


@pytest.mark.parametrize('shape', [(20000, 96, 16, 256, 5, 16), (6000, 768, 32, 256, 3, 16), (5000, 64, 8, 300, 7, 16), (4000, 32, 8, 256, 11, 24)])
def test_single_list_model_vs_oracle_at_gpu_test_shapes(shape):
    from helpers import tie_aware_rows
    N, D, M, Ks, seed, Mconn = shape
    g, t = _host_built(N, D, M, Ks, seed, threads=8, Mconn=Mconn)
    for k, ef in ((10, 64), (10, 256), (100, 128)):
        ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, k, ef, with_counts=True, with_ties=True)
        ml, md, mf, mh, mn = O.single_list_walk(g, t, k, ef)
        v = np.array(tie_aware_rows(ml, md, ol, od))
        clean = ties == 0
        assert (v[clean] == 'exact').all()
        assert np.array_equal(mh[clean], hops[clean]) and np.array_equal(mn[clean], nbrs[clean])
        assert (v[~clean] == 'diff').sum() <= 1


# ---- flagged single-list walk (filters / deletions): scalar C model vs the oracle -------------------------
@pytest.mark.parametrize('mode', ['filter', 'deleted'])
def test_flagged_model_vs_oracle_on_fixtures(golden, mode):
    from helpers import tie_aware_rows
    g = golden.oracle_graph(deleted=(mode == 'deleted'))
    fl = golden.allow if mode == 'filter' else None
    t = golden.query_tables_oracle()
    ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, golden.k, golden.ef, filter_labels=fl, with_counts=True, with_ties=True)
    ml, md, mf, mh, mn, peak = O.flagged_walk(g, t, golden.k, golden.ef, filter_labels=fl, cap=512)
    assert np.array_equal(mf, found)
    v = np.array(tie_aware_rows(ml, md, ol, od))
    clean = ties == 0
    assert (v[clean] == 'exact').all()
    assert np.array_equal(mh[clean], hops[clean]) and np.array_equal(mn[clean], nbrs[clean])
    assert (v[~clean] == 'diff').sum() <= (3 if golden.name == 'ties_k16' else 0)


def test_flagged_model_capacity_rule_and_overflow_flag():
    """The host's capacity rule (launch_search: 1.3*ef/s + 48) against the peak list size the walk reaches under
    random filters, and the overflow flag (found = -1) when the capacity is too small on purpose."""
    from helpers import tie_aware_rows
    g, t = _host_built(20000, 64, 8, 256, 21, threads=8)
    rng = np.random.default_rng(4)
    labels = g.labels()
    for ef, s in ((64, 0.9), (64, 0.5), (64, 0.3), (128, 0.5)):
        allow = labels[rng.random(g.n) < s]
        need = ef / (len(allow) / g.n) * 1.3 + 48
        cap = next(c for c in (64, 128, 256, 512) if c >= need and c >= ef)
        ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, 10, ef, filter_labels=allow, with_counts=True, with_ties=True)
        ml, md, mf, mh, mn, peak = O.flagged_walk(g, t, 10, ef, filter_labels=allow, cap=cap)
        assert (mf == found).all() and peak.max() <= cap          # no query overflows under a random filter
        v = np.array(tie_aware_rows(ml, md, ol, od))
        clean = ties == 0
        assert (v[clean] == 'exact').all() and np.array_equal(mh[clean], hops[clean])
        assert (v[~clean] == 'diff').sum() <= 1
    allow = labels[rng.random(g.n) < 0.3]
    ml, md, mf, mh, mn, peak = O.flagged_walk(g, t, 10, 64, filter_labels=allow, cap=128)   # needs ~ 64/0.3 = 213 entries
    assert (mf == -1).mean() > 0.9


# ---- two-list walk (hnsw_walk4f: admitted list P + traversed-only list N): scalar C model vs the oracle ----------
@pytest.mark.parametrize('mode', ['filter', 'deleted', 'plain'])
def test_two_list_model_vs_oracle_on_fixtures(golden, mode):
    from helpers import tie_aware_rows
    g = golden.oracle_graph(deleted=(mode == 'deleted'))
    fl = golden.allow if mode == 'filter' else None
    t = golden.query_tables_oracle()
    ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, golden.k, golden.ef, filter_labels=fl, with_counts=True, with_ties=True)
    ml, md, mf, mh, mn, peak = O.two_list_walk(g, t, golden.k, golden.ef, filter_labels=fl, cap_n=256)
    assert np.array_equal(mf, found)
    v = np.array(tie_aware_rows(ml, md, ol, od))
    clean = ties == 0
    assert (v[clean] == 'exact').all()
    assert np.array_equal(mh[clean], hops[clean]) and np.array_equal(mn[clean], nbrs[clean])
    assert (v[~clean] == 'diff').sum() <= (3 if golden.name == 'ties_k16' else 0)


def walk4f_en_for(ef, s):
    """launch_walk4f's capacity rule for the traversed-only list (walk_flagged4.cu: walk4f_en_for)."""
    s = min(1.0, max(1e-4, s))
    need = ef * (1 - s) / s + 6.5 * np.sqrt(ef * (1 - s)) / s + 4.0
    return next((en for en in (2, 3, 4, 5, 6, 8) if en * 32 >= need), 0)


def test_two_list_model_capacity_rule_and_overflow_flag():
    """N's capacity rule against the peak size N reaches under random filters; a query is flagged (found = -1) only
    when an entry that could still be expanded fell off N -- every unflagged row is the oracle's."""
    from helpers import tie_aware_rows
    g, t = _host_built(20000, 64, 8, 256, 21, threads=8)
    rng = np.random.default_rng(4)
    labels = g.labels()
    for ef, s in ((64, 0.9), (64, 0.5), (64, 0.45), (128, 0.6), (10, 0.5), (100, 0.7)):
        allow = labels[rng.random(g.n) < s]
        en = walk4f_en_for(ef, len(allow) / g.n)
        assert en > 0
        ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, 10, ef, filter_labels=allow, with_counts=True, with_ties=True)
        ml, md, mf, mh, mn, peak = O.two_list_walk(g, t, 10, ef, filter_labels=allow, cap_n=32 * en)
        assert (mf == found).all(), (ef, s, int((mf == -1).sum()), int(peak.max()))     # no query overflows under a random filter
        v = np.array(tie_aware_rows(ml, md, ol, od))
        clean = ties == 0
        assert (v[clean] == 'exact').all() and np.array_equal(mh[clean], hops[clean]) and np.array_equal(mn[clean], nbrs[clean])
        assert (v[~clean] == 'diff').sum() <= 1
    # a capacity that is too small on purpose: flagged rows are re-run by the host, the others must still be right
    allow = labels[rng.random(g.n) < 0.3]
    ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t, 10, 64, filter_labels=allow, with_counts=True, with_ties=True)
    for cap in (64, 128, 160):
        ml, md, mf, mh, mn, peak = O.two_list_walk(g, t, 10, 64, filter_labels=allow, cap_n=cap)
        ok = mf >= 0
        v = np.array(tie_aware_rows(ml[ok], md[ok], ol[ok], od[ok]))
        clean = ties[ok] == 0
        assert (v[clean] == 'exact').all() and np.array_equal(mh[ok][clean], hops[ok][clean])
        if cap == 64:
            assert (mf == -1).mean() > 0.9       # needs ~ 64 * 0.7 / 0.3 = 150 entries
    # deletions: every tenth node deleted (N stays tiny), and a filter that admits nothing near the entry point
    st_labels = labels[rng.random(g.n) < 0.02]
    ol, od, found, (hops, nbrs, _), ties = O.hnsw_search(g, t[:64], 5, 16, filter_labels=st_labels, with_counts=True, with_ties=True)
    ml, md, mf, mh, mn, peak = O.two_list_walk(g, t[:64], 5, 16, filter_labels=st_labels, cap_n=100000)
    assert np.array_equal(mf, found)
    v = np.array(tie_aware_rows(ml, md, ol, od))
    assert (v[ties == 0] == 'exact').all() and np.array_equal(mh[ties == 0], hops[ties == 0])
