"""PQIndex: exhaustive ADC linear scan with the reference's interface
(annlite/core/index/pq_index.py + flat_index.py storage), scan + top-k fused on the GPU (K2)."""
from typing import List, Optional

import numpy as np

from ..codec.pq import PQCodec


class PQIndex:
    def __init__(self, dim: int, pq_codec: PQCodec, initial_size: Optional[int] = None, expand_step_size: int = 10240,
                 **kwargs):
        assert pq_codec is not None
        self._dense_dim = dim
        self._pq_codec = pq_codec
        self.dim = pq_codec.n_subvectors
        self.dtype = pq_codec.code_dtype
        self.expand_step_size = expand_step_size
        self.initial_size = initial_size or expand_step_size
        self._capacity = self.initial_size
        self._size = 0
        self._data = np.zeros((self.initial_size, self.dim), dtype=self.dtype)
        self._dirty = True

    @property
    def capacity(self):
        return self._capacity

    @property
    def size(self):
        return self._size

    def add_with_ids(self, x: np.ndarray, ids: List[int]):
        codes = self._pq_codec.encode(np.ascontiguousarray(x, dtype=np.float32))
        while max(ids) >= self._capacity:
            self._data = np.concatenate((self._data, np.zeros((self.expand_step_size, self.dim), dtype=self.dtype)), axis=0)
            self._capacity += self.expand_step_size
        self._data[ids, :] = codes
        self._size += len(x)
        self._dirty = True

    def _sync(self, indices=None):
        e = self._pq_codec.engine
        if indices is not None:
            e.set_codes(self._data[indices])
            self._dirty = True
        elif self._dirty:
            e.set_codes(self._data)   # the reference scans the whole capacity (pq_index.py:39-40)
            self._dirty = False
        return e

    def search(self, x: np.ndarray, limit: int = 10, indices: Optional[np.ndarray] = None):
        """pq_index.py:29-56: (dists[limit] squared-L2, ids[limit]) for one query."""
        assert x.shape[-1] == self._pq_codec.dim, \
            f'the query embedding dimension does not match with index dimension: {x.shape[-1]} vs {self.dim}'
        d, i = self.search_batch(np.asarray(x, dtype=np.float32).reshape(1, -1), limit, indices)
        return d[0], i[0]

    def search_batch(self, x: np.ndarray, limit: int = 10, indices: Optional[np.ndarray] = None):
        e = self._sync(indices)
        # PQIndex builds its table with precompute_adc => the squared-L2 form whatever the codec metric
        from ... import pq_bind
        tables = pq_bind.batch_precompute_adc_table(np.ascontiguousarray(x, dtype=np.float32), self._pq_codec.d_subvector,
                                                    self._pq_codec.n_clusters, self._pq_codec.codebooks)
        ids, dists = e.scan_topk(tables=tables, k=limit)
        if indices is not None:
            ids = np.asarray(indices)[ids]
        return dists.astype(np.float64), ids   # the reference hands back float64-widened fp32 values

    def delete(self, ids: List[int]):
        raise RuntimeError(f'the deletion operation is not allowed for {self.__class__.__name__}!')

    def update_with_ids(self, x: np.ndarray, ids: List[int], **kwargs):
        self._data[ids, :] = self._pq_codec.encode(np.ascontiguousarray(x, dtype=np.float32))
        self._dirty = True

    def reset(self, capacity: Optional[int] = None):
        self._size = 0
        self._capacity = capacity or self.initial_size
        self._data = np.zeros((self._capacity, self.dim), dtype=self.dtype)
        self._dirty = True
