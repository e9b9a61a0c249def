"""Writes tests/golden/ref_codec_cosine.pkl: a PQCodec pickled by the REFERENCE's own class
(annlite/core/codec/base.py:26-27 ``dump`` = pickle protocol 4), so that the product's ``PQCodec.load`` can be
tested against a file an existing AnnLite workspace would hold.  Needs /root/reference and oracle/_ref
(``python oracle/build_ref.py``); run from the repo root:  python oracle/make_codec_fixture.py"""
import glob
import importlib.util
import os
import pickle
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/annlite'


def reference_pqcodec():
    pkg = types.ModuleType('annlite')            # annlite/__init__.py pulls in docarray: import the modules bare
    pkg.__path__ = [REF]
    sys.modules['annlite'] = pkg
    so = glob.glob(os.path.join(HERE, '_ref', 'pq_bind*.so'))[0]
    spec = importlib.util.spec_from_file_location('annlite.pq_bind', so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules['annlite.pq_bind'] = mod
    pkg.pq_bind = mod
    from annlite.core.codec.pq import PQCodec
    from annlite.enums import Metric
    return PQCodec, Metric


if __name__ == '__main__':
    PQCodec, Metric = reference_pqcodec()
    x = np.random.default_rng(0).standard_normal((300, 16)).astype(np.float32)
    codec = PQCodec(dim=16, n_subvectors=4, n_clusters=16, metric=Metric.COSINE, n_init=1)
    codec.fit(x, iter=5)
    out = os.path.join(HERE, '..', 'tests', 'golden')
    with open(os.path.join(out, 'ref_codec_cosine.pkl'), 'wb') as f:
        pickle.dump(codec, f, protocol=4)
    np.save(os.path.join(out, 'ref_codec_cosine_codebook.npy'), codec.codebooks)
    print('wrote ref_codec_cosine.pkl', os.path.getsize(os.path.join(out, 'ref_codec_cosine.pkl')), 'bytes')
