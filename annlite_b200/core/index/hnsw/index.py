"""HnswIndex with the reference's interface (annlite/core/index/hnsw/index.py) on the GPU backend.

`search` keeps the reference's one-query contract (returns row 0, sqrt for EUCLIDEAN); the hot
calls go through `annlite_b200.hnsw_bind.Index`.  Differences, all on the fast side:
* the query is not PQ-encoded (the reference encodes it and PQLookup then ignores the result);
* tables are built on the device inside the same call (K1 -> K3), normalising twice for COSINE as
  pre_process + get_dist_mat do;
* `search_batch` exposes the batched path the reference only has at the binding level.
"""
import math
import os
from pathlib import Path
from typing import List, Optional, Union

import numpy as np

from ....enums import ExpandMode, Metric
from ....hnsw_bind import Index


class HnswIndex:
    def __init__(self, dim: int, dtype=np.float32, metric: Metric = Metric.COSINE, ef_construction: int = 200,
                 ef_search: int = 50, max_connection: int = 16, pq_codec=None,
                 index_file: Optional[Union[str, Path]] = None, initial_size: Optional[int] = None,
                 expand_step_size: int = 10240, expand_mode: ExpandMode = ExpandMode.STEP, device: int = 0,
                 bruteforce_filter_below: Optional[int] = None, **kwargs):
        assert expand_step_size > 0
        if isinstance(metric, str):
            metric = Metric.from_string(metric)
        self.initial_size = initial_size or expand_step_size
        self.expand_step_size, self.expand_mode = expand_step_size, expand_mode
        self.dim, self.dtype, self.metric = dim, np.dtype(dtype), metric
        self._capacity = self.initial_size
        self.ef_construction, self.ef_search, self.max_connection = ef_construction, ef_search, max_connection
        self.pq_codec = pq_codec
        self.index_file = index_file
        self.device = device
        # opt-in: filters that admit at most this many ids are answered by an exact ADC scan over those ids
        # instead of the filtered graph walk (the reference's own TODO, hnsw/index.py:152); None = reference behaviour
        self.bruteforce_filter_below = bruteforce_filter_below
        if pq_codec is None:
            raise NotImplementedError('annlite_b200.HnswIndex is the PQ-encoded HNSW backend; pass a PQCodec '
                                      '(the float HNSW path is out of scope, SURVEY.md section 2 row 8)')
        self._init_hnsw_index()

    # ---- construction ---------------------------------------------------------------------------
    def _init_hnsw_index(self):
        self._index = Index(space=self.space_name, dim=self.dim, device=self.device)
        self._set_backend_pq = False
        if self.index_file:
            if not os.path.exists(self.index_file):
                raise FileNotFoundError(f'index path: {self.index_file} does not exist')
            self.load(self.index_file)
        elif self.pq_codec.is_trained:
            self._index.init_index(max_elements=self.capacity, ef_construction=self.ef_construction,
                                   M=self.max_connection, pq_codec=self.pq_codec)
            self._set_backend_pq = True
        else:
            self._index.init_index(max_elements=self.capacity, ef_construction=self.ef_construction,
                                   M=self.max_connection, pq_codec=None)
        self._index.set_ef(self.ef_search)

    def _ensure_backend(self):
        if not self.pq_codec.is_trained:
            raise RuntimeError('Please train the PQ before using HNSW quantization backend')
        if not self._set_backend_pq:
            self._index.loadPQ(self.pq_codec)
            self._set_backend_pq = True

    def load(self, index_file: Union[str, Path]):
        self._ensure_backend()
        self._index.load_index(str(index_file))
        self._capacity = max(self._capacity, self._index.max_elements)

    def dump(self, index_file: Union[str, Path]):
        self._index.save_index(str(index_file))

    def _prep(self, x):
        x = np.asarray(x)
        if x.ndim == 1:
            x = x.reshape((1, -1))
        if x.dtype != np.float32:
            x = x.astype(np.float32)
        return np.ascontiguousarray(x)

    @property
    def _normalize_rounds(self):
        return 2 if self.metric == Metric.COSINE else 0   # hnsw/index.py:28-29 + pq.py:309-310

    # ---- BaseIndex surface ------------------------------------------------------------------------
    def add_with_ids(self, x: 'np.ndarray', ids: List[int], num_threads: int = -1):
        """hnsw/index.py:125-137.  Codes and per-row tables come from the device in chunks."""
        self._ensure_backend()
        x = self._prep(x)
        max_id = int(max(ids)) + 1
        if max_id > self.capacity:
            expand_steps = math.ceil(max_id / self.expand_step_size)
            self._expand_capacity(expand_steps * self.expand_step_size)
        if self.metric == Metric.COSINE:   # pre_process's pass; the library adds get_dist_mat's own pass
            from ....math import l2_normalize
            x = np.ascontiguousarray(l2_normalize(x), dtype=np.float32)
        self._index.add_vectors(x, ids=np.asarray(ids, dtype=np.uint64), num_threads=num_threads)

    def search(self, query: 'np.ndarray', limit: int = 10, indices: Optional['np.ndarray'] = None):
        """hnsw/index.py:140-167: one query -> (dists[limit], ids[limit]); EUCLIDEAN gets sqrt."""
        dists, ids = self.search_batch(self._prep(query)[:1], limit=limit, indices=indices)
        return dists[0], ids[0]

    def search_batch(self, queries, limit: int = 10, indices=None, out_ids=None, out_dists=None):
        """Batched search: (B, dim) -> (dists (B, limit), ids (B, limit)).  `queries` may be a numpy
        array (host) or a torch CUDA tensor; device outputs can be supplied to stay on the GPU."""
        self._ensure_backend()
        if isinstance(queries, np.ndarray) or not hasattr(queries, 'data_ptr'):
            queries = self._prep(queries)
        if indices is not None and len(indices) < limit:
            limit = len(indices)
        if limit <= 0:      # an empty candidate list: nothing can be returned (the reference would ask hnswlib for k=0)
            B = queries.shape[0]
            return np.empty((B, 0), dtype=np.float32), np.empty((B, 0), dtype=np.uint64)
        from ...._lib import MAX_EF
        if max(self.ef_search, limit) > MAX_EF:
            raise ValueError(f'limit / ef_search above {MAX_EF} is not supported by the GPU walk (ANNB_MAX_EF); '
                             f'use PQIndex (exhaustive scan) or lower limit={limit} / ef_search={self.ef_search}')
        self._index.set_ef(max(self.ef_search, limit))
        if (indices is not None and self.bruteforce_filter_below is not None
                and len(indices) <= self.bruteforce_filter_below and out_ids is None):
            ids, dists = self._index._e.scan_subset(queries, indices, k=limit, normalize=self._normalize_rounds)
            return (np.sqrt(dists) if self.metric == Metric.EUCLIDEAN else dists), ids
        ids, dists = self._index.knn_query_vectors(queries, k=limit, normalize=self._normalize_rounds,
                                                   filters=indices, out_labels=out_ids, out_dists=out_dists)
        if self.metric == Metric.EUCLIDEAN:
            if isinstance(dists, np.ndarray):
                dists = np.sqrt(dists)
            else:
                dists.sqrt_()
        return dists, ids

    def search_batch_submit(self, queries, limit: int = 10, indices=None):
        """Streamed batched search for serving loops: returns a ticket at once, two batches can be in flight so
        the transfers (and the filter upload) of one overlap the graph walk of the other.  `indices` = the allowed
        ids, as in `search` (hnsw/index.py:140-167)."""
        self._ensure_backend()
        if isinstance(queries, np.ndarray) or not hasattr(queries, 'data_ptr'):
            queries = self._prep(queries)
        if indices is not None and len(indices) < limit:
            raise ValueError('fewer allowed ids than `limit`: use search_batch (it shrinks the limit like the reference)')
        from ...._lib import MAX_EF
        if max(self.ef_search, limit) > MAX_EF:
            raise ValueError(f'limit / ef_search above {MAX_EF} is not supported by the GPU walk (ANNB_MAX_EF)')
        self._index.set_ef(max(self.ef_search, limit))
        return self._index.knn_query_submit(queries, k=limit, normalize=self._normalize_rounds, filters=indices)

    def search_batch_wait(self, ticket):
        ids, dists = self._index.knn_query_wait(ticket)
        if self.metric == Metric.EUCLIDEAN:
            dists = np.sqrt(dists) if isinstance(dists, np.ndarray) else dists.sqrt_()
        return dists, ids

    def delete(self, ids: List[int]):
        for i in ids:
            self._index.mark_deleted(i)

    def update_with_ids(self, x: 'np.ndarray', ids: List[int], **kwargs):
        raise RuntimeError(f'the update operation is not allowed for {self.__class__.__name__}!')

    def _expand_capacity(self, new_capacity: int):
        self._capacity = new_capacity
        self._index.resize_index(new_capacity)

    def reset(self, capacity: Optional[int] = None):
        self._capacity = capacity or self.initial_size
        self._init_hnsw_index()

    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def size(self):
        return self._index.element_count

    @property
    def space_name(self):
        if self.metric == Metric.EUCLIDEAN:
            return 'l2'
        elif self.metric == Metric.INNER_PRODUCT:
            return 'ip'
        return 'cosine'

    @property
    def pq_enable(self):
        return self.pq_codec is not None

    @property
    def normalization_enable(self):
        return self.metric == Metric.COSINE
