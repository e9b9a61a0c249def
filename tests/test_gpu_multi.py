"""GPU, 2 ranks, NCCL: the range-sharded search step (annlite_b200.sharded.ShardedEngine) -- walk of the rank's
shard, ONE all-gather of the packed (B,k) {fp32, u64} results on the Engine's own stream, merge kernel -- against
G oracle graphs (one per shard, searchKnn restatement) merged by the reference's rule
(annlite/container.py:130-138, host statement merge_topk_host).  Skipped with fewer than 2 GPUs: run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q` (output committed under profiles/)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
        import oracle as O
        from annlite_b200.engine import Engine
        from annlite_b200.sharded import ShardedEngine, merge_topk_host, shard_range
        N, D, M, B, k, ef = 40_000, 128, 8, 512, 10, 64
        rng = np.random.default_rng(11)
        X = rng.standard_normal((N, D)).astype(np.float32)
        cb = np.stack([X[rng.choice(N, 256, replace=False), m * 16:(m + 1) * 16] for m in range(M)]).astype(np.float32)
        Qs = rng.standard_normal((6, B, D)).astype(np.float32)
        lo, hi = shard_range(N, rank, world)
        e = Engine(D, M, 256, 'euclidean', device=rank)
        e.set_codebook(cb)
        e.init_graph(hi - lo, M=16, ef_construction=100)
        e.add_items(X[lo:hi], np.arange(lo, hi, dtype=np.uint64))          # labels = global ids
        g = O.Graph.from_state(e.get_graph(), M, 256)
        se = ShardedEngine(e, B, k)
        Qd = torch.from_numpy(Qs).cuda()
        hl = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(6)]
        hd = [torch.empty((B, k), dtype=torch.float32).pin_memory() for _ in range(6)]
        tickets = []
        for i in range(6):                                                  # two batches in flight
            if len(tickets) == 2:
                se.wait(tickets.pop(0))
            tickets.append(se.submit(Qd[i], ef, host_labels=hl[i], host_dists=hd[i]))
        for t in tickets:
            se.wait(t)
        e.sync()
        bad = 0
        for i in range(6):
            tq = O.adc_table(Qs[i], cb, 'euclidean')
            ol, od, found = O.hnsw_search(g, tq, k, ef)
            both = [None] * world
            dist.all_gather_object(both, (ol, od))
            L = np.stack([b[0] for b in both])
            Dd = np.stack([b[1] for b in both])
            ml, md = merge_topk_host(L, Dd, k)
            got_l, got_d = hl[i].numpy().view(np.uint64), hd[i].numpy()
            bad += int(not (np.array_equal(got_l, ml) and np.array_equal(got_d.view(np.uint32), md.view(np.uint32))))
        # host-free data path: nothing but kernels, one all-gather and the result copies were enqueued per step
        launches = e.launch_count
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad, launches))
    except Exception as ex:  # noqa: BLE001
        import traceback
        q.put((rank, -1, traceback.format_exc()))
        raise


def test_sharded_step_two_ranks_nccl_matches_merged_oracle_graphs():
    torch = pytest.importorskip('torch')
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=600) for _ in procs]
    [p.join(60) for p in procs]
    for rank, bad, info in res:
        assert bad == 0, (rank, bad, info)
