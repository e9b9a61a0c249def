"""Multi-GPU layout of the search path: one process per GPU (`torch.distributed`).

Two layouts:

* **replicate** -- every rank holds the whole index and serves its own slice of the query batch.
  No data-path collective at all; results are gathered only if the caller wants them in one place.
* **shard**     -- base vectors are range-partitioned by id, rank r owns [r*N/G, (r+1)*N/G) with its
  own HNSW graph over that slice (labels stay global ids), every rank sees all queries, and the
  per-shard top-k lists are merged with ONE all-gather of (B, k) {fp32 dist, u64 label} followed by
  the k-way merge kernel (annb_merge_topk).  The merge rule is the reference's own
  (annlite/container.py:130-138: concatenate per-cell results, sort by distance, keep `limit`),
  made deterministic with (dist, label) ordering.

The class is transport-agnostic so that the host logic is testable on CPU with the gloo backend:
`local_search` and `merge` are injected (GPU: Engine.search / Engine.merge_topk on device tensors).
"""
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous id range of shard `rank` (the same split bench.py --mode shard builds)."""
    return rank * n // world, (rank + 1) * n // world


def query_slice(b: int, rank: int, world: int) -> Tuple[int, int]:
    """Slice of a B-query batch served by `rank` in the replicate layout."""
    return rank * b // world, (rank + 1) * b // world


def merge_topk_host(labels_gbk: np.ndarray, dists_gbk: np.ndarray, k: int):
    """Host statement of the merge rule (container.py:130-138) with the (dist, label) tie order.
    Used by the CPU tests as the specification of annb_merge_topk; missing entries carry
    label == UINT64_MAX and are ignored."""
    G, B, kk = labels_gbk.shape
    out_l = np.full((B, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    out_d = np.full((B, k), np.inf, dtype=np.float32)
    for b in range(B):
        l = labels_gbk[:, b, :].reshape(-1)
        d = dists_gbk[:, b, :].reshape(-1)
        keep = l != np.iinfo(np.uint64).max
        l, d = l[keep], d[keep]
        order = np.lexsort((l, d))[:k]
        out_l[b, :len(order)] = l[order]
        out_d[b, :len(order)] = d[order]
    return out_l, out_d


class ShardedSearcher:
    """search() over a range-sharded index: local walk -> all-gather -> merge."""

    def __init__(self, local_search: Callable, merge: Optional[Callable] = None, group=None):
        self.local_search = local_search
        self.merge = merge
        self.group = group

    def search(self, queries, k: int):
        import torch
        import torch.distributed as dist
        labels, dists = self.local_search(queries, k)          # (B,k) on this rank's shard
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return labels, dists
        as_t = lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        tl, td = as_t(labels), as_t(dists)
        if tl.dtype != torch.int64:                            # uint64 labels travel as their int64 bit pattern
            tl = tl.view(torch.int64)
        # concatenated along dim 0 (the form every backend accepts), viewed as (G, B, k) afterwards
        gl = torch.empty((world * tl.shape[0],) + tuple(tl.shape[1:]), dtype=tl.dtype, device=tl.device)
        gd = torch.empty((world * td.shape[0],) + tuple(td.shape[1:]), dtype=td.dtype, device=td.device)
        dist.all_gather_into_tensor(gl, tl.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gd, td.contiguous(), group=self.group)
        gl = gl.view((world,) + tuple(tl.shape))
        gd = gd.view((world,) + tuple(td.shape))
        if self.merge is not None:                             # device merge kernel
            return self.merge(gl, gd, k)
        ml, md = merge_topk_host(gl.cpu().numpy().view(np.uint64), gd.cpu().numpy(), k)
        return ml, md


class ShardedEngine:
    """The sharded search step on one GPU of G (one process per GPU, NCCL): walk of this rank's shard, ONE
    all-gather of the packed (B,k) {fp32 dist, u64 label} results, merge -- all enqueued on one stream of the
    Engine, no host synchronisation in between, two batches in flight (the walk of batch i+1 hides the gather and
    merge of batch i).  Merge rule = annlite/container.py:130-138 with (dist, label) order.

    Every rank must call submit() in the same order (the all-gathers of one process group are matched by order)."""

    def __init__(self, engine, B, k, group=None):
        import torch
        import torch.distributed as dist
        self.e, self.B, self.k, self.group = engine, int(B), int(k), group
        self.world = dist.get_world_size(group)
        self.off_l = (self.B * self.k * 4 + 7) // 8 * 8
        self.stride = self.off_l + self.B * self.k * 8
        dev = torch.device('cuda', engine.device)
        self.packed = [torch.zeros(self.stride, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.gathered = [torch.zeros(self.world * self.stride, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.out_l = [torch.empty((self.B, self.k), dtype=torch.int64, device=dev) for _ in range(2)]
        self.out_d = [torch.empty((self.B, self.k), dtype=torch.float32, device=dev) for _ in range(2)]
        self.streams = [torch.cuda.ExternalStream(engine.lane_stream(i), device=dev) for i in range(2)]

    def views(self, lane):
        n = self.B * self.k
        p = self.packed[lane]
        import torch
        return (p[self.off_l:self.off_l + n * 8].view(torch.int64).view(self.B, self.k),
                p[:n * 4].view(torch.float32).view(self.B, self.k))

    def submit(self, queries, ef, normalize=0, host_labels=None, host_dists=None):
        """queries: (B, D) device tensor or pinned host array, the same on every rank.  Returns a ticket; the
        merged result of that ticket is in .result(ticket) (device) and, if given, in the pinned host tensors."""
        import torch
        import torch.distributed as dist
        lane = self.e.next_lane
        lv, dv = self.views(lane)
        t = self.e.search_submit(queries, lv, dv, k=self.k, ef=ef, normalize=normalize)
        assert (t & 1) == lane
        with torch.cuda.stream(self.streams[lane]):
            dist.all_gather_into_tensor(self.gathered[lane], self.packed[lane], group=self.group)
        self.e.merge_topk_packed(self.gathered[lane], self.world, self.B, self.k, self.stride, self.off_l,
                                 self.out_l[lane], self.out_d[lane], lane)
        if host_labels is not None:
            with torch.cuda.stream(self.streams[lane]):
                host_labels.copy_(self.out_l[lane], non_blocking=True)
                host_dists.copy_(self.out_d[lane], non_blocking=True)
        return t

    def wait(self, ticket):
        self.e.search_wait(ticket)

    def result(self, ticket):
        lane = ticket & 1
        return self.out_l[lane], self.out_d[lane]
