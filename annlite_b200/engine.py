"""`Engine`: one annb_index_t handle (one GPU, one CUDA stream) with numpy / torch friendly methods.

This is the thin host layer between the reference-shaped Python classes (hnsw_bind.Index,
pq_bind, PQCodec, HnswIndex, PQIndex, AnnLite) and the C ABI.  PyTorch tensors are accepted
as device buffers (their data_ptr is passed through); results are numpy (host) unless device
output tensors are supplied.
"""
import ctypes as C

import numpy as np

from . import _lib as L

_METRIC = {'l2': L.METRIC_L2, 'euclidean': L.METRIC_L2, 'ip': L.METRIC_IP, 'inner_product': L.METRIC_IP,
           'cosine': L.METRIC_COSINE}


def metric_code(metric):
    if isinstance(metric, str):
        return _METRIC[metric.lower()]
    name = getattr(metric, 'name', None)
    if name is not None:
        return _METRIC[name.lower()]
    return int(metric)


class Engine:
    def __init__(self, dim, n_subvectors, n_clusters, metric='euclidean', device=0):
        self._lib = L.load()
        self._h = C.c_void_p()
        self.dim, self.M, self.Ks = int(dim), int(n_subvectors), int(n_clusters)
        self.metric = metric_code(metric)
        self.device = int(device)
        L.check(self._lib.annb_create(self.device, self.metric, self.dim, self.M, self.Ks, C.byref(self._h)))
        self.code_dtype = np.uint8 if self.Ks <= 256 else np.uint16
        self.has_codebook = False

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.annb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- codebook / tables --------------------------------------------------------------
    def set_codebook(self, cb):
        if isinstance(cb, np.ndarray):
            cb = np.ascontiguousarray(cb, dtype=np.float32)
        if tuple(cb.shape) != (self.M, self.Ks, self.dim // self.M):
            raise AttributeError('PQ class returning the codebook with wrong dimension')
        p, sp, keep = L.as_ptr(cb)
        L.check(self._lib.annb_set_codebook(self._h, p, sp))
        self.has_codebook = True

    def adc_table(self, queries, normalize=0, out=None):
        q = self._as_f32(queries)
        B = q.shape[0]
        if out is None:
            out = np.empty((B, self.M, self.Ks), dtype=np.float32)
        qp, qs, _k1 = L.as_ptr(q)
        op, os_, _k2 = L.as_ptr(out)
        L.check(self._lib.annb_adc_table(self._h, qp, qs, B, int(normalize), op, os_))
        return out

    # ---- exhaustive scan ----------------------------------------------------------------
    def set_codes(self, codes):
        if isinstance(codes, np.ndarray):
            codes = np.ascontiguousarray(codes, dtype=self.code_dtype)
        p, sp, _k = L.as_ptr(codes)
        L.check(self._lib.annb_set_codes(self._h, p, sp, codes.shape[0]))
        self.n_codes = codes.shape[0]

    def scan(self, table, out=None):
        t = self._as_f32(table)
        if out is None:
            out = np.empty(self.n_codes, dtype=np.float32)
        tp, ts, _k1 = L.as_ptr(t)
        op, os_, _k2 = L.as_ptr(out)
        L.check(self._lib.annb_scan(self._h, tp, ts, op, os_))
        return out

    def scan_topk(self, queries=None, tables=None, k=10, out_ids=None, out_dists=None):
        src = self._as_f32(queries if queries is not None else tables)
        B = src.shape[0]
        if out_ids is None:
            out_ids = np.empty((B, k), dtype=np.int64)
            out_dists = np.empty((B, k), dtype=np.float32)
        sp_, ss, _k0 = L.as_ptr(src)
        ip, is_, _k1 = L.as_ptr(out_ids)
        dp, ds_, _k2 = L.as_ptr(out_dists)
        assert is_ == ds_
        L.check(self._lib.annb_scan_topk(self._h, sp_ if queries is not None else None,
                                         sp_ if queries is None else None, ss, B, int(k), ip, dp, is_))
        return out_ids, out_dists

    # ---- graph --------------------------------------------------------------------------
    def init_graph(self, max_elements, M=16, ef_construction=200, random_seed=100):
        L.check(self._lib.annb_init_graph(self._h, int(max_elements), int(M), int(ef_construction), int(random_seed)))

    def load_index(self, path, max_elements=0):
        L.check(self._lib.annb_load_index(self._h, str(path).encode(), int(max_elements)))

    def save_index(self, path):
        L.check(self._lib.annb_save_index(self._h, str(path).encode()))

    def set_graph(self, st):
        """Adopt the reference's pickle dict (Index.__getstate__()[0])."""
        l0 = np.ascontiguousarray(np.asarray(st['data_level0']).view(np.uint8))
        ll = np.ascontiguousarray(np.asarray(st['link_lists']).view(np.uint8))
        if ll.size == 0:
            ll = np.zeros(8, dtype=np.uint8)
        lv = np.ascontiguousarray(np.asarray(st['element_levels'], dtype=np.int32))
        L.check(self._lib.annb_set_graph(
            self._h, l0.ctypes.data, int(l0.nbytes), int(st['size_data_per_element']), int(st['offset_data']),
            int(st['label_offset']), ll.ctypes.data, int(np.asarray(st['link_lists']).nbytes), lv.ctypes.data, int(lv.size),
            int(st['size_links_per_element']), int(st['cur_element_count']),
            int(st['max_elements']), int(st['max_level']), int(st['enterpoint_node']) & 0xFFFFFFFF, int(st['max_M']),
            int(st['max_M0']), int(st['M']), int(st['ef_construction']), float(st['mult'])))

    def graph_info(self):
        n, mx, lb, spe = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_uint64()
        ml, ep = C.c_int32(), C.c_uint32()
        mM, mM0, M, efc, mult = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_double()
        L.check(self._lib.annb_graph_info(self._h, C.byref(n), C.byref(mx), C.byref(spe), C.byref(lb), C.byref(ml),
                                          C.byref(ep), C.byref(mM), C.byref(mM0), C.byref(M), C.byref(efc), C.byref(mult)))
        return dict(cur_element_count=n.value, max_elements=mx.value, size_data_per_element=spe.value,
                    link_lists_bytes=lb.value, max_level=ml.value, enterpoint_node=ep.value, max_M=mM.value,
                    max_M0=mM0.value, M=M.value, ef_construction=efc.value, mult=mult.value)

    def get_graph(self):
        """The pickle-dict arrays of the reference (hnsw_bindings.cpp:623-671)."""
        info = self.graph_info()
        n = info['cur_element_count']
        l0 = np.zeros(n * info['size_data_per_element'], dtype=np.uint8)
        ll = np.zeros(max(info['link_lists_bytes'], 1), dtype=np.uint8)
        lv = np.zeros(max(n, 1), dtype=np.int32)
        L.check(self._lib.annb_get_graph(self._h, l0.ctypes.data, ll.ctypes.data, lv.ctypes.data))
        crow = self.M * np.dtype(self.code_dtype).itemsize
        info.update(data_level0=l0, link_lists=ll[:info['link_lists_bytes']], element_levels=lv[:n],
                    offset_level0=0, offset_data=4 + 4 * info['max_M0'],
                    label_offset=4 + 4 * info['max_M0'] + crow, size_links_per_element=4 + 4 * info['max_M'])
        return info

    def add_items(self, vectors, labels, codes=None, num_threads=-1):
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        lab = np.ascontiguousarray(labels, dtype=np.uint64)
        assert v.ndim == 2 and v.shape[1] == self.dim and lab.shape[0] == v.shape[0]
        cp = None
        if codes is not None:
            codes = np.ascontiguousarray(codes, dtype=self.code_dtype)
            cp = codes.ctypes.data
        L.check(self._lib.annb_add_items(self._h, v.ctypes.data, cp, lab.ctypes.data, v.shape[0], int(num_threads)))

    def add_items_with_tables(self, codes, tables, labels, num_threads=-1):
        codes = np.ascontiguousarray(codes, dtype=self.code_dtype)
        tables = np.ascontiguousarray(tables, dtype=np.float32)
        lab = np.ascontiguousarray(labels, dtype=np.uint64)
        L.check(self._lib.annb_add_items_with_tables(self._h, codes.ctypes.data, tables.ctypes.data, lab.ctypes.data,
                                                     codes.shape[0], int(num_threads)))

    def encode(self, vectors):
        v = self._as_f32(vectors)
        out = np.empty((v.shape[0], self.M), dtype=self.code_dtype)
        vp, vs, _k = L.as_ptr(v)
        L.check(self._lib.annb_encode(self._h, vp, vs, v.shape[0], out.ctypes.data, L.HOST))
        return out

    def resize_index(self, n):
        L.check(self._lib.annb_resize_index(self._h, int(n)))

    def mark_deleted(self, label):
        L.check(self._lib.annb_mark_deleted(self._h, int(label)))

    def unmark_deleted(self, label):
        L.check(self._lib.annb_unmark_deleted(self._h, int(label)))

    @property
    def element_count(self):
        n = C.c_int64()
        L.check(self._lib.annb_element_count(self._h, C.byref(n)))
        return n.value

    def get_labels(self):
        n = self.element_count
        out = np.empty(n, dtype=np.uint64)
        L.check(self._lib.annb_get_labels(self._h, out.ctypes.data, n))
        return out

    def get_codes(self, labels):
        lab = np.ascontiguousarray(labels, dtype=np.uint64)
        out = np.empty((lab.shape[0], self.M), dtype=self.code_dtype)
        L.check(self._lib.annb_get_codes(self._h, lab.ctypes.data, lab.shape[0], out.ctypes.data))
        return out

    # ---- search -------------------------------------------------------------------------
    def search(self, queries=None, tables=None, k=10, ef=50, normalize=0, filter_labels=None, with_stats=False,
               out_labels=None, out_dists=None):
        src = self._as_f32(queries if queries is not None else tables)
        B = src.shape[0]
        if out_labels is None:
            out_labels = np.empty((B, k), dtype=np.uint64)
            out_dists = np.empty((B, k), dtype=np.float32)
        stats = np.zeros((B, 3), dtype=np.int64) if with_stats else None
        sp_, ss, _k0 = L.as_ptr(src)
        lp, ls, _k1 = L.as_ptr(out_labels)
        dp, ds_, _k2 = L.as_ptr(out_dists)
        assert ls == ds_
        fp, fs, fn, _k3 = None, L.HOST, 0, None
        if filter_labels is not None:
            if isinstance(filter_labels, np.ndarray) or not hasattr(filter_labels, 'data_ptr'):
                filter_labels = np.ascontiguousarray(filter_labels, dtype=np.uint64)
            fp, fs, _k3 = L.as_ptr(filter_labels)
            fn = int(filter_labels.shape[0])
            if fn == 0:
                fp = np.zeros(1, dtype=np.uint64).ctypes.data
        stp = None
        if stats is not None:
            if ls == L.DEVICE:
                raise ValueError('with_stats needs host outputs')
            stp = stats.ctypes.data
        rc = self._lib.annb_search(self._h, sp_ if queries is not None else None, sp_ if queries is None else None, ss, B,
                                   int(normalize), int(k), int(ef), fp, fs, fn, lp, dp, ls, stp)
        L.check(rc)
        if with_stats:
            return out_labels, out_dists, stats
        return out_labels, out_dists

    def scan_subset(self, queries, subset_labels, k=10, normalize=0):
        """Exact ADC over the indexed nodes whose labels are in `subset_labels` (brute-force route for very
        selective filters).  Returns (labels uint64 (B,k), dists fp32 (B,k))."""
        q = self._as_f32(queries)
        sub = np.ascontiguousarray(subset_labels, dtype=np.uint64)
        B = q.shape[0]
        labels = np.empty((B, k), dtype=np.uint64)
        dists = np.empty((B, k), dtype=np.float32)
        qp, qs, _k = L.as_ptr(q)
        L.check(self._lib.annb_scan_subset(self._h, qp, qs, B, int(normalize), int(k), sub.ctypes.data if sub.size else None,
                                           sub.shape[0], labels.ctypes.data, dists.ctypes.data))
        return labels, dists

    def search_submit(self, queries, out_labels, out_dists, k=10, ef=50, normalize=0, filter_labels=None):
        """Streaming form: enqueue one batch (host numpy or device torch buffers) and return a ticket; up
        to two batches are in flight.  The buffers must stay alive and untouched until `search_wait`.
        `filter_labels` = the allowed labels (knn_query_with_filter), host or device."""
        q = self._as_f32(queries)
        qp, qs, k0 = L.as_ptr(q)
        lp, ls, k1 = L.as_ptr(out_labels)
        dp, ds_, k2 = L.as_ptr(out_dists)
        assert ls == ds_
        t = C.c_int()
        k3 = None
        if filter_labels is None:
            L.check(self._lib.annb_search_submit(self._h, qp, qs, q.shape[0], int(normalize), int(k), int(ef), lp, dp, ls,
                                                 C.byref(t)))
        else:
            if isinstance(filter_labels, np.ndarray) or not hasattr(filter_labels, 'data_ptr'):
                filter_labels = np.ascontiguousarray(filter_labels, dtype=np.uint64)
            fn = int(filter_labels.shape[0])
            if fn == 0:
                filter_labels = np.zeros(1, dtype=np.uint64)
            fp, fs, k3 = L.as_ptr(filter_labels)
            L.check(self._lib.annb_search_submit_filtered(self._h, qp, qs, q.shape[0], int(normalize), int(k), int(ef), fp, fs, fn,
                                                          lp, dp, ls, C.byref(t)))
        self._inflight = getattr(self, '_inflight', {})
        self._inflight[t.value] = (k0, k1, k2, k3)      # keep the buffers alive
        self._next_ticket = t.value + 1
        return t.value

    @property
    def next_lane(self):
        """Internal lane (0/1 = stream) the next search_submit will use: tickets alternate."""
        return getattr(self, '_next_ticket', 0) & 1

    def search_wait(self, ticket):
        try:
            L.check(self._lib.annb_search_wait(self._h, int(ticket)))
        finally:
            getattr(self, '_inflight', {}).pop(ticket, None)

    def merge_topk(self, labels_gbk, dists_gbk, out_labels, out_dists):
        G, B, k = labels_gbk.shape
        L.check(self._lib.annb_merge_topk(self._h, labels_gbk.data_ptr(), dists_gbk.data_ptr(), G, B, k,
                                          out_labels.data_ptr(), out_dists.data_ptr()))

    def merge_topk_packed(self, gathered, G, B, k, rank_stride_bytes, labels_offset_bytes, out_labels, out_dists, lane):
        """Merge G all-gathered packed shard results ([dists | labels] per rank) on lane `lane`'s stream."""
        L.check(self._lib.annb_merge_topk_packed(self._h, gathered.data_ptr(), int(G), int(B), int(k), int(rank_stride_bytes),
                                                 int(labels_offset_bytes), out_labels.data_ptr(), out_dists.data_ptr(), int(lane)))

    def lane_stream(self, lane):
        s = C.c_uint64()
        L.check(self._lib.annb_lane_stream(self._h, int(lane), C.byref(s)))
        return s.value

    # ---- misc ---------------------------------------------------------------------------
    def sync(self):
        L.check(self._lib.annb_sync(self._h))

    @property
    def stream(self):
        s = C.c_uint64()
        L.check(self._lib.annb_stream(self._h, C.byref(s)))
        return s.value

    def last_kernel_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        L.check(self._lib.annb_last_kernel_ms(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(table_ms=a.value, search_ms=b.value, scan_ms=c.value)

    @property
    def launch_count(self):
        n = C.c_int64()
        L.check(self._lib.annb_launch_count(self._h, C.byref(n)))
        return n.value

    @property
    def fallback_count(self):
        n = C.c_int64()
        L.check(self._lib.annb_fallback_count(self._h, C.byref(n)))
        return n.value

    @property
    def fallback_queries(self):
        n = C.c_int64()
        L.check(self._lib.annb_fallback_queries(self._h, C.byref(n)))
        return n.value

    @property
    def sync_counts(self):
        """(full re-derivations of the device graph, incremental patches)"""
        a, b = C.c_int64(), C.c_int64()
        L.check(self._lib.annb_sync_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_option(self, name, value):
        L.check(self._lib.annb_set_option(self._h, name.encode(), int(value)))

    @staticmethod
    def _as_f32(x):
        if isinstance(x, np.ndarray):
            x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim == 1:
                x = x[None]
            return x
        if hasattr(x, 'data_ptr'):
            import torch
            if x.dtype != torch.float32:
                x = x.float()
            return x.contiguous()
        return np.ascontiguousarray(np.asarray(x, dtype=np.float32))
