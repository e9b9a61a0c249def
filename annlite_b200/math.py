"""Host-side helpers with the reference's semantics (annlite/math.py)."""
import numpy as np


def l2_normalize(x: 'np.ndarray', eps: float = np.finfo(np.float32).eps):
    """annlite/math.py:6-18.  Host version (used when the caller wants numpy in / numpy out);
    the search path normalises on the device (annb_adc_table / annb_search `normalize`)."""
    norms = np.einsum('ij,ij->i', x, x)
    np.sqrt(norms, norms)
    norms[norms < 10 * eps] = 1.0
    return x / norms[:, np.newaxis]


def top_k(values: 'np.ndarray', k: int, descending: bool = False):
    """annlite/math.py:94-120 (used only by host-side merges of tiny lists)."""
    if descending:
        values = -values
    if k >= values.shape[1]:
        idx = values.argsort(axis=1)[:, :k]
        values = np.take_along_axis(values, idx, axis=1)
    else:
        idx_ps = values.argpartition(kth=k, axis=1)[:, :k]
        values = np.take_along_axis(values, idx_ps, axis=1)
        idx_fs = values.argsort(axis=1)
        idx = np.take_along_axis(idx_ps, idx_fs, axis=1)
        values = np.take_along_axis(values, idx_fs, axis=1)
    if descending:
        values = -values
    return values, idx
