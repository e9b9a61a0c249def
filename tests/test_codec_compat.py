"""CPU: a PQCodec pickled by the reference's own class (fixture made by oracle/make_codec_fixture.py) loads
through annlite_b200's PQCodec.load with the same state; next-row f3 (existing AnnLite workspaces)."""
import os
import pickle

import numpy as np
import pytest

from annlite_b200.core.codec.pq import PQCodec
from annlite_b200.enums import Metric

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_reference_pickle_loads_with_the_same_state():
    c = PQCodec.load(os.path.join(G, 'ref_codec_cosine.pkl'))
    assert isinstance(c, PQCodec) and c.is_trained and c.metric is Metric.COSINE and c.normalize_input
    assert (c.dim, c.n_subvectors, c.n_clusters, c.d_subvector) == (16, 4, 16, 4) and c.code_dtype == np.uint8
    assert np.array_equal(c.get_codebook(), np.load(os.path.join(G, 'ref_codec_cosine_codebook.npy')))
    assert c.get_subspace_splitting() == (4, 16, 4) and len(c.kmeans) == 4


def test_own_dump_load_round_trip(tmp_path):
    c = PQCodec.load(os.path.join(G, 'ref_codec_cosine.pkl'))
    p = tmp_path / 'codec.pkl'
    c.dump(p)
    d = PQCodec.load(p)
    assert d.metric is Metric.COSINE and np.array_equal(d.codebooks, c.codebooks) and d._engine is None


def test_foreign_reference_classes_are_refused(tmp_path, monkeypatch):
    import sys
    import types
    mod = types.ModuleType('annlite.executor')            # a pickle naming some other reference class
    mod.AnnLiteIndexer = type('AnnLiteIndexer', (), {'__module__': 'annlite.executor'})
    monkeypatch.setitem(sys.modules, 'annlite', types.ModuleType('annlite'))
    monkeypatch.setitem(sys.modules, 'annlite.executor', mod)
    p = tmp_path / 'x.pkl'
    p.write_bytes(pickle.dumps(mod.AnnLiteIndexer(), protocol=4))
    with pytest.raises(pickle.UnpicklingError, match='no counterpart'):
        PQCodec.load(p)
