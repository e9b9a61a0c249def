"""CPU: AnnLite keeps the reference's workspace layout (annlite/index.py:573-640, :679-710, :769-777):
`parameters-<md5>/pq_codec.params` and `snapshot-<md5>/<time>-SNAPSHOT/cell_0.hnsw`, the md5 taken over the same
text -- so a directory written by the reference opens here for the codec + graph part.  The graph object is
stubbed (it needs the GPU); the codec is the fixture pickled by the reference's own class."""
import hashlib
import os
import shutil

import pytest

import annlite_b200.index as mod
from annlite_b200.enums import Metric

G = os.path.join(os.path.dirname(__file__), 'golden')


class FakeGraph:
    made = []

    def __init__(self, dim, metric=None, pq_codec=None, **kw):
        self.dim, self.metric, self.codec, self.kw, self.loaded, self.size = dim, metric, pq_codec, kw, None, 0
        FakeGraph.made.append(self)

    def dump(self, path):
        with open(path, 'wb') as f:
            f.write(b'graph')

    def load(self, path):
        self.loaded, self.size = str(path), 123


@pytest.fixture
def fake(monkeypatch):
    FakeGraph.made.clear()
    monkeypatch.setattr(mod, 'HnswIndex', FakeGraph)
    return FakeGraph


def ref_hash(n_dim, metric, n_subvectors):
    text = f'n_dim: {n_dim} metric: {metric} n_cells: 1 n_components: None n_subvectors: {n_subvectors}'
    return hashlib.md5(text.encode()).hexdigest()


def test_paths_follow_the_reference(tmp_path, fake):
    a = mod.AnnLite(16, metric='cosine', n_subvectors=4, data_path=tmp_path)
    h = ref_hash(16, 'COSINE', 4)
    assert a.params_hash == h
    assert a.model_path == tmp_path / f'parameters-{h}' and a._pq_codec_path.name == 'pq_codec.params'
    assert a.index_path.parent == tmp_path / f'snapshot-{h}' and a.index_path.name.endswith('-SNAPSHOT')
    assert '#' in a.index_path.name and a.snapshot_path is None and not a.is_trained and not fake.made
    assert mod.AnnLite(16, metric='euclidean', n_subvectors=4, data_path=tmp_path).params_hash == ref_hash(16, 'EUCLIDEAN', 4)


def test_opens_a_workspace_written_by_the_reference(tmp_path, fake):
    h = ref_hash(16, 'COSINE', 4)
    (tmp_path / f'parameters-{h}').mkdir()
    shutil.copy(os.path.join(G, 'ref_codec_cosine.pkl'), tmp_path / f'parameters-{h}' / 'pq_codec.params')
    for stamp in ('2024-01-01#00:00:00', '2024-03-01#12:30:00'):      # the later snapshot wins (index.py:628-637)
        d = tmp_path / f'snapshot-{h}' / f'{stamp}-SNAPSHOT'
        d.mkdir(parents=True)
        (d / 'cell_0.hnsw').write_bytes(b'x')
    a = mod.AnnLite(16, metric='cosine', n_subvectors=4, n_clusters=16, data_path=tmp_path)
    assert a.is_trained and a._pq_codec.metric is Metric.COSINE and len(fake.made) == 1
    g = fake.made[0]
    assert g.codec is a._pq_codec and g.loaded.endswith('2024-03-01#12:30:00-SNAPSHOT/cell_0.hnsw') and a.index_size == 123


def test_dump_writes_the_same_layout_and_restore_reads_it_back(tmp_path, fake):
    h = ref_hash(16, 'COSINE', 4)
    (tmp_path / f'parameters-{h}').mkdir()
    shutil.copy(os.path.join(G, 'ref_codec_cosine.pkl'), tmp_path / f'parameters-{h}' / 'pq_codec.params')
    a = mod.AnnLite(16, metric='cosine', n_subvectors=4, n_clusters=16, data_path=tmp_path)
    with pytest.raises(FileNotFoundError):
        a.restore()
    snap = a.dump()
    assert snap.parent.name == f'snapshot-{h}' and (snap / 'cell_0.hnsw').read_bytes() == b'graph'
    assert a.snapshot_path == snap
    a.restore()
    assert fake.made[0].loaded == str(snap / 'cell_0.hnsw')
    b = mod.AnnLite(16, metric='cosine', n_subvectors=4, n_clusters=16, data_path=tmp_path)   # reopen: restores by itself
    assert b.index_size == 123


def test_small_surface_calls_that_need_no_gpu(tmp_path, fake):
    import numpy as np
    h = ref_hash(16, 'COSINE', 4)
    (tmp_path / f'parameters-{h}').mkdir()
    shutil.copy(os.path.join(G, 'ref_codec_cosine.pkl'), tmp_path / f'parameters-{h}' / 'pq_codec.params')
    a = mod.AnnLite(16, metric='cosine', n_subvectors=4, n_clusters=16, data_path=tmp_path)
    x = np.random.default_rng(0).standard_normal((5, 16)).astype(np.float32)
    assert a.encode(x) is x                                   # the reference's one-cell quirk (index.py:557)
    codes = np.random.default_rng(1).integers(0, 16, (5, 4)).astype(np.uint8)
    rec = a.decode(codes)
    assert rec.shape == (5, 16) and np.array_equal(rec[:, :4], a._pq_codec.codebooks[0][codes[:, 0]])
    assert a.vec_index(0) is a.cell_indexes[0] and a.total_docs == a.index_size
    with pytest.raises(IndexError):
        a.vec_index(1)
    with pytest.raises(NotImplementedError):
        a.backup('somewhere', token='t')
    assert a.backup().name.endswith('-SNAPSHOT')
