"""CPU: re-adding existing labels (hnswalg.h:958-1096 updatePoint / repairConnectionsForUpdate -- what
AnnLite.update does through add_with_ids, annlite/container.py:343-347) reproduces the compiled reference's
graph byte for byte (single-threaded), including the un-delete of a re-added deleted label."""
import numpy as np
import pytest

import oracle as O
from annlite_b200.engine import Engine
from oracle import ref_driver as R

pytestmark = pytest.mark.skipif(not R.available(), reason='oracle/_ref not built')


@pytest.mark.parametrize('metric,seed', [('euclidean', 1), ('cosine', 2), ('inner_product', 3)])
def test_update_existing_labels_matches_reference(metric, seed):
    rng = np.random.default_rng(seed)
    N, D, M, Ks = 1500, 32, 4, 64
    X = rng.standard_normal((N, D)).astype(np.float32)
    Xn = R.l2_normalize(X).astype(np.float32) if metric == 'cosine' else X
    ds = D // M
    cb = np.stack([Xn[rng.choice(N, Ks, replace=False), m * ds:(m + 1) * ds] for m in range(M)]).astype(np.float32)
    codec = R.RefCodec(cb, metric)
    labels = rng.permutation(N).astype(np.uint64) + 3
    ref = R.RefHnswIndex(codec, metric, capacity=N + 10, ef_construction=60, ef_search=32, max_connection=8)
    ref.add_with_ids(X, labels, num_threads=1)

    e = Engine(D, M, Ks, metric, device=-1)
    e.init_graph(N + 10, M=8, ef_construction=60)
    Xp = ref._pre(X)
    e.add_items_with_tables(codec.encode(Xp), codec.get_dist_mat(Xp), labels, num_threads=1)
    assert np.array_equal(e.get_graph()['data_level0'], np.asarray(ref.state()['data_level0']).view(np.uint8))

    # 1) update 40 stored points with new vectors, mixed with 5 brand-new labels in the same call
    upd = rng.choice(N, 40, replace=False)
    new_lab = np.concatenate([labels[upd], np.arange(10 ** 6, 10 ** 6 + 5, dtype=np.uint64)])
    Y = rng.standard_normal((45, D)).astype(np.float32)
    order = rng.permutation(45)
    new_lab, Y = new_lab[order], Y[order]
    ref.add_with_ids(Y, new_lab, num_threads=1)
    Yp = ref._pre(Y)
    e.add_items_with_tables(codec.encode(Yp), codec.get_dist_mat(Yp), new_lab, num_threads=1)
    a, b = e.get_graph(), ref.state()
    assert a['cur_element_count'] == b['cur_element_count'] == N + 5
    assert np.array_equal(a['data_level0'], np.asarray(b['data_level0']).view(np.uint8))
    assert np.array_equal(a['link_lists'], np.asarray(b['link_lists']).view(np.uint8))
    assert a['enterpoint_node'] == b['enterpoint_node'] and a['max_level'] == b['max_level']

    # 2) delete a few labels, then re-add two of them: the mark is cleared and the point re-linked
    for l in labels[upd[:6]]:
        ref._index.mark_deleted(int(l))
        e.mark_deleted(int(l))
    Z = rng.standard_normal((2, D)).astype(np.float32)
    back = labels[upd[:2]]
    ref.add_with_ids(Z, back, num_threads=1)
    Zp = ref._pre(Z)
    e.add_items_with_tables(codec.encode(Zp), codec.get_dist_mat(Zp), back, num_threads=1)
    a, b = e.get_graph(), ref.state()
    assert np.array_equal(a['data_level0'], np.asarray(b['data_level0']).view(np.uint8))
    assert np.array_equal(a['link_lists'], np.asarray(b['link_lists']).view(np.uint8))
    g = O.Graph.from_state(a, M, Ks)
    assert g.links0()[2].sum() == 4

    # 3) searches over the updated graph agree (oracle on our graph == reference on its own)
    Q = rng.standard_normal((30, D)).astype(np.float32)
    T = codec.get_dist_mat(ref._pre(Q))
    rl, rd = ref.knn_query(Q, 5, num_threads=1, tables=T)
    ol, od, _ = O.hnsw_search(g, T, 5, 32)
    assert np.array_equal(rl, ol) and np.array_equal(rd.view(np.uint32), od.view(np.uint32))


def test_update_only_entry_point_graph():
    codec_cb = np.random.default_rng(0).standard_normal((2, 8, 4)).astype(np.float32)
    codec = R.RefCodec(codec_cb, 'euclidean')
    ref = R.RefHnswIndex(codec, 'euclidean', capacity=8, ef_construction=10, max_connection=4)
    x = np.random.default_rng(1).standard_normal((1, 8)).astype(np.float32)
    ref.add_with_ids(x, np.array([7], dtype=np.uint64), num_threads=1)
    ref.add_with_ids(x * 2, np.array([7], dtype=np.uint64), num_threads=1)     # single element: early return (:965)
    e = Engine(8, 2, 8, 'euclidean', device=-1)
    e.init_graph(8, M=4, ef_construction=10)
    for v in (x, x * 2):
        e.add_items_with_tables(codec.encode(v), codec.get_dist_mat(v), np.array([7], dtype=np.uint64), num_threads=1)
    assert np.array_equal(e.get_graph()['data_level0'], np.asarray(ref.state()['data_level0']).view(np.uint8))
    assert e.element_count == 1
