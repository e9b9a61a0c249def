"""CPU: host-side graph container and insertion algorithm (no GPU: host-only handle, device=-1).

* single-threaded insertion reproduces the reference-built graph byte for byte (fixtures hold the
  graph the compiled reference built single-threaded from the same codes and tables);
* hnswlib save/load format round-trips and is parsed identically by the oracle's reader;
* pickle-state import/export is lossless;
* compute entry points refuse to run on a host-only handle (no CPU fallback).
"""
import os

import numpy as np
import pytest

import oracle as O
from annlite_b200 import _lib as L
from annlite_b200.engine import Engine


def insert_tables(fx):
    X = fx.X
    if fx.metric == 'cosine':  # pre_process normalises, get_dist_mat normalises again
        X = O.l2_normalize(X).astype(np.float32)
    return O.adc_table(X, fx.cb, fx.metric)


def host_engine(fx):
    return Engine(fx.M * fx.ds, fx.M, fx.Ks, fx.metric, device=-1)


def test_single_thread_build_is_byte_identical(golden):
    e = host_engine(golden)
    st = golden.state
    e.init_graph(st['max_elements'], M=st['M'], ef_construction=st['ef_construction'], random_seed=100)
    e.add_items_with_tables(golden.codes, insert_tables(golden), golden.labels, num_threads=1)
    got = e.get_graph()
    assert got['cur_element_count'] == st['cur_element_count']
    assert got['max_level'] == st['max_level'] and got['enterpoint_node'] == st['enterpoint_node']
    assert np.array_equal(got['element_levels'], st['element_levels'][:len(got['element_levels'])])
    assert np.array_equal(got['data_level0'], st['data_level0'])
    assert np.array_equal(got['link_lists'], st['link_lists'])
    assert got['mult'] == st['mult']


def test_incremental_build_matches_one_shot(golden):
    e = host_engine(golden)
    st = golden.state
    T = insert_tables(golden)
    e.init_graph(st['max_elements'], M=st['M'], ef_construction=st['ef_construction'])
    n = len(golden.codes)
    cuts = [0, 1, 7, n // 3, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        e.add_items_with_tables(golden.codes[a:b], T[a:b], golden.labels[a:b], num_threads=1)
    got = e.get_graph()
    assert np.array_equal(got['data_level0'], st['data_level0'])
    assert np.array_equal(got['link_lists'], st['link_lists'])


def test_multithreaded_build_is_a_valid_graph(golden):
    e = host_engine(golden)
    st = golden.state
    e.init_graph(st['max_elements'], M=st['M'], ef_construction=st['ef_construction'])
    e.add_items_with_tables(golden.codes, insert_tables(golden), golden.labels, num_threads=4)
    got = e.get_graph()
    n = got['cur_element_count']
    assert n == st['cur_element_count']
    g = O.Graph.from_state(got, golden.M, golden.Ks)
    cnt, lk, dele = g.links0()
    assert cnt.max() <= g.max_M0 and cnt.min() >= 1
    for i in range(0, n, 97):
        nb = lk[i, :cnt[i]]
        assert len(set(nb.tolist())) == len(nb) and i not in nb and nb.max() < n
    assert sorted(g.labels().tolist()) == sorted(golden.labels.tolist())
    # searching it with the oracle finds every query's own nearest code reasonably often
    t = golden.query_tables_oracle()
    l, d, found = O.hnsw_search(g, t, golden.k, golden.ef)
    assert (found == golden.k).all()
    g_ref = golden.oracle_graph()
    l_ref, d_ref, _ = O.hnsw_search(g_ref, t, golden.k, golden.ef)
    # same data, different (thread-order dependent) graph: mean best distance must be comparable
    assert abs(float(d[:, 0].mean()) - float(d_ref[:, 0].mean())) <= 0.25 * abs(float(d_ref[:, 0].mean())) + 1e-3


def test_save_load_roundtrip(golden, tmp_path):
    e = host_engine(golden)
    e.set_graph(golden.state)
    p = tmp_path / 'cell_0.hnsw'
    e.save_index(p)
    g = O.Graph.from_save_file(str(p), golden.M, golden.Ks)   # independent parser (oracle)
    assert np.array_equal(g.level0, golden.state['data_level0'])
    assert np.array_equal(g.links[:len(golden.state['link_lists'])], golden.state['link_lists'])
    assert g.enterpoint == golden.state['enterpoint_node'] and g.maxlevel == golden.state['max_level']
    e2 = host_engine(golden)
    e2.load_index(p)
    got = e2.get_graph()
    assert np.array_equal(got['data_level0'], golden.state['data_level0'])
    assert np.array_equal(got['link_lists'], golden.state['link_lists'])
    assert np.array_equal(got['element_levels'], golden.state['element_levels'][:got['cur_element_count']])
    assert np.array_equal(np.sort(e2.get_labels()), np.sort(golden.labels))
    # codes come back by label (Index.get_items)
    some = golden.labels[[0, 5, 17]]
    assert np.array_equal(e2.get_codes(some), golden.codes[[0, 5, 17]])


def test_load_rejects_garbage(golden, tmp_path):
    p = tmp_path / 'bad.hnsw'
    p.write_bytes(b'\x00' * 50)
    e = host_engine(golden)
    with pytest.raises(RuntimeError, match='corrupted|unsupported'):
        e.load_index(p)
    with pytest.raises(RuntimeError, match='Cannot open file'):
        e.load_index(tmp_path / 'missing.hnsw')


def test_mark_deleted_and_errors(golden):
    e = host_engine(golden)
    e.set_graph(golden.state)
    lab = int(golden.labels[3])
    e.mark_deleted(lab)
    with pytest.raises(RuntimeError, match='already deleted'):
        e.mark_deleted(lab)
    with pytest.raises(RuntimeError, match='Label not found'):
        e.mark_deleted(10 ** 15)
    got = e.get_graph()
    g = O.Graph.from_state(got, golden.M, golden.Ks)
    assert g.links0()[2].sum() == 1
    with pytest.raises(RuntimeError, match='Label not found'):
        e.get_codes(np.array([lab], dtype=np.uint64))   # deleted labels are hidden (hnswalg.h:853)
    e.unmark_deleted(lab)
    assert not O.Graph.from_state(e.get_graph(), golden.M, golden.Ks).links0()[2].any()


def test_capacity_and_resize(golden):
    e = host_engine(golden)
    T = insert_tables(golden)
    e.init_graph(10, M=golden.state['M'], ef_construction=50)
    with pytest.raises(RuntimeError, match='exceeds the specified limit'):
        e.add_items_with_tables(golden.codes[:11], T[:11], golden.labels[:11], num_threads=1)
    e.resize_index(64)
    e.add_items_with_tables(golden.codes[:11], T[:11], golden.labels[:11], num_threads=1)
    assert e.element_count == 11
    # re-adding a stored label updates it in place (hnswalg.h:1119-1131): the count does not grow
    e.add_items_with_tables(golden.codes[:1], T[:1], golden.labels[:1], num_threads=1)
    assert e.element_count == 11


def test_no_cpu_fallback_on_host_only_handle(golden):
    e = host_engine(golden)
    e.set_graph(golden.state)
    with pytest.raises(L.AnnbError) as ei:
        e.search(tables=golden.tables, k=golden.k, ef=golden.ef)
    assert ei.value.code == L.ENODEVICE
    with pytest.raises(L.AnnbError):
        e.set_codebook(golden.cb)
    with pytest.raises(L.AnnbError):
        e.adc_table(golden.Q)


# ---- hostile inputs: a graph from outside is checked before it can reach the GPU -------------------------
def _small_graph(tmp_path, Ks=16):
    import oracle as O
    rng = np.random.default_rng(0)
    X = rng.standard_normal((600, 16)).astype(np.float32)
    cb = np.stack([X[rng.choice(600, Ks, replace=False), m * 4:(m + 1) * 4] for m in range(4)]).astype(np.float32)
    e = Engine(16, 4, Ks, 'euclidean', device=-1)
    e.init_graph(600, M=8, ef_construction=40)
    e.add_items_with_tables(O.encode(X, cb), O.adc_table(X, cb), np.arange(600, dtype=np.uint64), num_threads=1)
    p = str(tmp_path / 'g.hnsw')
    e.save_index(p)
    return e, p, open(p, 'rb').read()


def test_corrupted_index_files_are_rejected_not_trusted(tmp_path):
    import struct
    e, p, good = _small_graph(tmp_path)
    spe = struct.unpack_from('<Q', good, 24)[0]

    def load(mut):
        q = str(tmp_path / 'bad.hnsw')
        open(q, 'wb').write(bytes(mut))
        Engine(16, 4, 16, 'euclidean', device=-1).load_index(q)

    load(good)                                                   # the untouched file loads
    cases = {}
    b = bytearray(good); b[8:16] = struct.pack('<Q', 10); b[16:24] = struct.pack('<Q', 1 << 40); cases['count beyond the file'] = b
    b = bytearray(good); b[8:16] = struct.pack('<Q', 3); cases['limit below count'] = None      # tolerated: clamped to the count
    b = bytearray(good); b[52:56] = struct.pack('<I', 100000); cases['entry point out of range'] = b
    b = bytearray(good); b[48:52] = struct.pack('<i', 40); cases['max level above every node'] = b
    b = bytearray(good); b[96 + 4:96 + 8] = struct.pack('<I', 0x7fffffff); cases['level-0 link out of range'] = b
    b = bytearray(good); b[96:98] = struct.pack('<H', 999); cases['level-0 count above maxM0'] = b
    b = bytearray(good); b[96 + spe - 12] = 200; cases['code not below Ks'] = b                 # Ks=16, u8 codes
    b = bytearray(good[:len(good) // 2]); cases['truncated'] = b
    b = bytearray(good); b[80:88] = struct.pack('<d', float('nan')); cases['mult is NaN'] = b
    for name, mut in cases.items():
        if mut is None:
            continue
        with pytest.raises(L.AnnbError):
            load(mut)


def test_set_graph_checks_the_state_it_is_given(tmp_path):
    e, p, good = _small_graph(tmp_path)
    st = e.get_graph()
    Engine(16, 4, 16, 'euclidean', device=-1).set_graph(st)      # the untouched state is accepted
    bad = dict(st); l0 = np.array(st['data_level0']).copy().view(np.uint8); l0[4:8] = np.frombuffer(np.uint32(123456).tobytes(), np.uint8); bad['data_level0'] = l0
    with pytest.raises(L.AnnbError):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)
    bad = dict(st); bad['enterpoint_node'] = 60000
    with pytest.raises(L.AnnbError):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)
    lv = np.array(st['element_levels']).copy(); lv[5] = -3; bad = dict(st); bad['element_levels'] = lv
    with pytest.raises(L.AnnbError):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)
    # arrays shorter than the counts claim: the extents travel with the pointers (include/annb.h)
    bad = dict(st); bad['link_lists'] = np.array(st['link_lists'])[:-8]
    with pytest.raises(L.AnnbError, match='link-list bytes'):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)
    bad = dict(st); bad['data_level0'] = np.array(st['data_level0'])[:-80]
    with pytest.raises(L.AnnbError, match='inconsistent'):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)
    bad = dict(st); bad['cur_element_count'] = 0            # an "empty" graph that still names an entry point
    with pytest.raises(L.AnnbError):
        Engine(16, 4, 16, 'euclidean', device=-1).set_graph(bad)


def test_codes_outside_the_codebook_are_refused(tmp_path):
    import oracle as O
    e = Engine(16, 4, 16, 'euclidean', device=-1)
    e.init_graph(10, M=8, ef_construction=40)
    codes = np.full((3, 4), 16, dtype=np.uint8)                  # valid codes are 0..15
    with pytest.raises(L.AnnbError, match='not below n_clusters'):
        e.add_items_with_tables(codes, np.zeros((3, 4, 16), np.float32), np.arange(3, dtype=np.uint64), num_threads=1)
    assert e.element_count == 0
