import numpy as np


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def tie_aware_rows(labels, dists, ref_labels, ref_dists):
    """Per-row verdict comparing a result against the oracle/reference.

    'exact'  : labels and fp32 distances identical
    'tie'    : distances bit-identical, labels differ only where equal distances make the order
               (or the choice at the k-th boundary) ambiguous -- the only freedom the GPU walk takes
    'diff'   : anything else
    """
    out = []
    k = labels.shape[1]
    for r in range(labels.shape[0]):
        if np.array_equal(labels[r], ref_labels[r]) and np.array_equal(bits(dists[r]), bits(ref_dists[r])):
            out.append('exact')
            continue
        if not np.array_equal(bits(dists[r]), bits(ref_dists[r])):
            out.append('diff')
            continue
        ok = True
        for p in np.nonzero(labels[r] != ref_labels[r])[0]:
            d = dists[r, p]
            tied = (p > 0 and dists[r, p - 1] == d) or (p + 1 < k and dists[r, p + 1] == d) or p == k - 1
            ok &= bool(tied)
        out.append('tie' if ok else 'diff')
    return out


def recall(pred, truth):
    """annlite/utils.py:52-71: |pred ∩ truth| / |truth| averaged over queries."""
    return float(np.mean([len(set(p.tolist()) & set(t.tolist())) / len(t) for p, t in zip(pred, truth)]))
