"""Driver for the compiled reference modules in ``oracle/_ref`` (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline legs
may import this.  It loads ``pq_bind`` / ``hnsw_bind`` (built by ``oracle/build_ref.py`` from
``/root/reference``) and drives them the way the reference's own Python layer does, without
importing the reference's Python package (which cannot travel to the GPU box and needs
docarray):

* ``RefCodec``      -- the duck-typed object ``Index._loadPQ`` expects
                       (bindings/hnsw_bindings.cpp:851-903): ``encode``, ``get_codebook``,
                       ``get_subspace_splitting``; plus ``get_dist_mat`` restating
                       annlite/core/codec/pq.py:293-325 on top of the *reference's* pq_bind.
* ``RefHnswIndex``  -- restates annlite/core/index/hnsw/index.py:20-48,125-167
                       (pre_process -> add_items / knn_query / knn_query_with_filter -> sqrt).
* ``ref_pq_linear_scan`` -- restates annlite/core/index/pq_index.py:29-56 +
                       annlite/math.py:94-120 on top of the reference's pq_bind.
"""
import importlib.util
import os
import sysconfig

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, '_ref')
_mods = {}


def available() -> bool:
    suffix = sysconfig.get_config_var('EXT_SUFFIX') or '.so'
    return all(os.path.exists(os.path.join(_REF, n + suffix)) for n in ('pq_bind', 'hnsw_bind'))


def _load(name):
    if name not in _mods:
        suffix = sysconfig.get_config_var('EXT_SUFFIX') or '.so'
        path = os.path.join(_REF, name + suffix)
        if not os.path.exists(path):
            raise ImportError(f'{path} missing: run `python oracle/build_ref.py` where /root/reference exists')
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def pq_bind():
    return _load('pq_bind')


def hnsw_bind():
    return _load('hnsw_bind')


def l2_normalize(x, eps=np.finfo(np.float32).eps):
    """annlite/math.py:6-18 (restated)."""
    norms = np.einsum('ij,ij->i', x, x)
    np.sqrt(norms, norms)
    norms[norms < 10 * eps] = 1.0
    return x / norms[:, np.newaxis]


class RefCodec:
    """Trained-PQ stand-in: the codebook is an input (training is out of scope)."""

    def __init__(self, codebooks: np.ndarray, metric: str = 'euclidean'):
        cb = np.ascontiguousarray(codebooks, dtype=np.float32)
        self.n_subvectors, self.n_clusters, self.d_subvector = cb.shape
        self.dim = self.n_subvectors * self.d_subvector
        self.codebooks = cb
        self.metric = metric
        self.is_trained = True
        self.code_dtype = (np.uint8 if self.n_clusters <= 2 ** 8
                           else (np.uint16 if self.n_clusters <= 2 ** 16 else np.uint32))

    # --- protocol required by Index._loadPQ ---------------------------------------
    def get_codebook(self):
        return self.codebooks

    def get_subspace_splitting(self):
        return (self.n_subvectors, self.n_clusters, self.d_subvector)

    def encode(self, x):
        """annlite/core/codec/pq.py:158-177 (scipy.cluster.vq.vq per subspace)."""
        from scipy.cluster.vq import vq
        x = np.ascontiguousarray(x, dtype=np.float32)
        codes = np.empty((x.shape[0], self.n_subvectors), dtype=self.code_dtype)
        ds = self.d_subvector
        for m in range(self.n_subvectors):
            codes[:, m], _ = vq(x[:, m * ds:(m + 1) * ds], self.codebooks[m])
        return codes

    # --- annlite/core/codec/pq.py:293-325 -------------------------------------------
    def get_dist_mat(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        if self.metric == 'cosine':
            x = l2_normalize(x)
        pb = pq_bind()
        if self.metric == 'euclidean':
            t = pb.batch_precompute_adc_table(x, self.d_subvector, self.n_clusters, self.codebooks)
        else:
            t = 1 / self.n_clusters - np.array(
                pb.batch_precompute_adc_table_ip(x, self.d_subvector, self.n_clusters, self.codebooks),
                dtype='float32')
        return np.ascontiguousarray(t, dtype='float32')

    def precompute_adc(self, q):
        """annlite/core/codec/pq.py:200-224 (single query, L2 only as in the reference)."""
        return np.asarray(pq_bind().precompute_adc_table(
            np.ascontiguousarray(q, dtype=np.float32), self.d_subvector, self.n_clusters, self.codebooks))


_SPACE = {'euclidean': 'l2', 'inner_product': 'ip', 'cosine': 'cosine'}


class RefHnswIndex:
    """annlite/core/index/hnsw/index.py restated over the compiled reference ``hnsw_bind.Index``."""

    def __init__(self, codec: RefCodec, metric='euclidean', capacity=10240, ef_construction=200,
                 ef_search=50, max_connection=16):
        self.codec, self.metric, self.ef_search = codec, metric, ef_search
        self._index = hnsw_bind().Index(space=_SPACE[metric], dim=codec.dim)
        self._index.init_index(max_elements=capacity, ef_construction=ef_construction,
                               M=max_connection, pq_codec=codec)
        self._index.set_ef(ef_search)

    def _pre(self, x):
        x = np.asarray(x, dtype=np.float32)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        if self.metric == 'cosine':
            x = l2_normalize(x)
        return x

    def add_with_ids(self, x, ids, num_threads=-1, batch=5000, codes=None):
        """hnsw/index.py:125-137; inserts in slices so the (n, M, Ks) tables stay small."""
        x = self._pre(x)
        ids = np.asarray(ids, dtype=np.uint64)
        for s in range(0, len(x), batch):
            xs = x[s:s + batch]
            tables = self.codec.get_dist_mat(xs)
            cs = self.codec.encode(xs) if codes is None else codes[s:s + batch]
            self._index.add_items(cs, ids=ids[s:s + batch], num_threads=num_threads, dtables=tables)

    def knn_query(self, x, k=10, num_threads=-1, indices=None, tables=None):
        """Batched native entry (hnsw_bindings.cpp:303-375 / :393-495); squared/raw distances."""
        x = self._pre(x)
        if tables is None:
            tables = self.codec.get_dist_mat(x)
        codes = np.zeros((len(x), self.codec.n_subvectors), dtype=self.codec.code_dtype)  # unused by PQLookup
        self._index.set_ef(max(self.ef_search, k))
        if indices is not None:
            return self._index.knn_query_with_filter(codes, filters=np.asarray(indices, dtype=np.uint64),
                                                     k=k, num_threads=num_threads, dtables=tables)
        return self._index.knn_query(codes, k=k, num_threads=num_threads, dtables=tables)

    def search(self, query, limit=10, indices=None):
        """hnsw/index.py:140-167: one query, encode included (as shipped), returns row 0."""
        x = self._pre(query)
        tables = self.codec.get_dist_mat(x)
        codes = self.codec.encode(x)
        self._index.set_ef(max(self.ef_search, limit))
        if indices is not None:
            if len(indices) < limit:
                limit = len(indices)
            ids, dists = self._index.knn_query_with_filter(codes, filters=indices, k=limit, dtables=tables)
        else:
            ids, dists = self._index.knn_query(codes, k=limit, dtables=tables)
        if self.metric == 'euclidean':
            dists = np.sqrt(dists)
        return dists[0], ids[0]

    def state(self):
        """Index.__getstate__()[0] (hnsw_bindings.cpp:549-689): the exported graph."""
        return self._index.__getstate__()[0]


def top_k(values, k):
    """annlite/math.py:94-120 restated (ascending)."""
    if k >= values.shape[1]:
        idx = values.argsort(axis=1)[:, :k]
        values = np.take_along_axis(values, idx, axis=1)
    else:
        idx_ps = values.argpartition(kth=k, axis=1)[:, :k]
        values = np.take_along_axis(values, idx_ps, axis=1)
        idx_fs = values.argsort(axis=1)
        idx = np.take_along_axis(idx_ps, idx_fs, axis=1)
        values = np.take_along_axis(values, idx_fs, axis=1)
    return values, idx


def ref_pq_linear_scan(codec: RefCodec, codes: np.ndarray, q: np.ndarray, limit=10):
    """annlite/core/index/pq_index.py:29-56: single-query exhaustive ADC + top_k."""
    table = codec.precompute_adc(q)
    dists = np.asarray(pq_bind().dist_pqcodes_to_codebooks(table, codes))
    d, i = top_k(dists[None, :], limit)
    return d[0], i[0]
