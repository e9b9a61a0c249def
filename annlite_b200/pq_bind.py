"""Drop-in for `annlite.pq_bind` (bindings/pq_bindings.pyx): same function names, argument
order and return shapes, computed by the CUDA kernels K1 / K2 through the C ABI.

The reference functions are stateless (the codebook travels with every call); an Engine per
(codebook buffer, metric) is cached here so repeated calls do not re-upload it.
"""
import hashlib

import numpy as np

from .engine import Engine

_cache = {}
_MAX_CACHE = 8


def _engine(codebooks, ip: bool, device=0):
    cb = np.ascontiguousarray(codebooks, dtype=np.float32)
    # keyed by the codebook's CONTENT (a few hundred KB: hashing it costs microseconds): an in-place edit or a new
    # array at a recycled address must never meet a stale copy on the GPU
    key = (hashlib.blake2b(cb.tobytes(), digest_size=16).digest(), cb.shape, bool(ip), device)
    e = _cache.get(key)
    if e is None:
        if len(_cache) >= _MAX_CACHE:
            _cache.pop(next(iter(_cache))).close()
        M, Ks, ds = cb.shape
        e = Engine(M * ds, M, Ks, 'inner_product' if ip else 'euclidean', device=device)
        e.set_codebook(cb)
        _cache[key] = e
    return e


def precompute_adc_table(query, d_subvector, n_clusters, codebooks):
    """pq_bindings.pyx:85-145: (D,) query -> (M, Ks) fp32 squared-L2 table."""
    q = np.ascontiguousarray(query, dtype=np.float32).reshape(1, -1)
    return _engine(codebooks, False).adc_table(q)[0]


def batch_precompute_adc_table(queries, d_subvector, n_clusters, codebooks):
    """pq_bindings.pyx:149-210: (N, D) -> (N, M, Ks) fp32 squared-L2 tables."""
    return _engine(codebooks, False).adc_table(np.ascontiguousarray(queries, dtype=np.float32))


def batch_precompute_adc_table_ip(queries, d_subvector, n_clusters, codebooks):
    """pq_bindings.pyx:214-274: (N, D) -> (N, M, Ks) raw inner products (the `1/Ks - .` step is the
    caller's, pq.py:316-322)."""
    e = _engine(codebooks, True)
    e.set_option('ip_raw', 1)           # T = 0 - ip: negation is exact, so the raw products come back bit-exact
    try:
        t = e.adc_table(np.ascontiguousarray(queries, dtype=np.float32))
    finally:
        e.set_option('ip_raw', 0)
    return np.negative(t)


def dist_pqcodes_to_codebooks(adtable, pq_codes):
    """pq_bindings.pyx:52-80: one (M, Ks) table against (N, M) codes -> N distances.
    Returns a float32 ndarray (the reference returns a Python list of the same values)."""
    table = np.ascontiguousarray(adtable, dtype=np.float32)
    codes = np.ascontiguousarray(pq_codes)
    M, Ks = table.shape
    key = ('scan', M, Ks)
    e = _cache.get(key)
    if e is None:
        e = Engine(M, M, Ks, 'euclidean')   # geometry only: no codebook needed for a scan
        _cache[key] = e
    e.set_codes(codes.astype(e.code_dtype, copy=False))
    return e.scan(table)
