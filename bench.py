#!/usr/bin/env python
"""bench.py -- queries/sec of the PQ-HNSW search path (BASELINE.json configs[1]: 1M x 128d fp32,
M=8 PQ, HNSW M=16 ef_construction=200, ef=64, k=10) on N B200s of one node.

A "step" is one pass of the hot path (K1 ADC tables -> K3 HNSW walk -> top-k) over one batch of
synthetic queries.  Prints ONE JSON line (rank 0):
  value     : whole-job QPS with the query batch already resident in HBM (device in, device out)
  e2e       : same metric through the host-buffer C-ABI call (H2D of queries + K1 + K3 + D2H of results)
  roofline  : dominant kernel (K3 walk): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline : the C oracle port timed on the host cores on a bounded sample of the same batch
`--impl reference` times the reference's own CPU path (oracle/_ref: compiled from /root/reference)
on the same config: batched get_dist_mat + knn_query with all host threads.

Multi-GPU (`torchrun ... bench.py --gpus N`): the index is replicated (1M x 384 B records = 384 MB
fits every GPU) and every rank serves its own query batch -> no data-path collective, weak
scaling.  `--mode shard` range-shards the base vectors instead (one graph per rank, labels =
global ids), replicates the queries and merges per-shard top-k with one NCCL all-gather + the
merge kernel.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
CACHE = os.path.join(ROOT, '.index_cache')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n', '--base-n', dest='n', type=int, default=1_000_000)   # use --base-n under torchrun (its parser treats --n as ambiguous)
    ap.add_argument('--dim', type=int, default=128)
    ap.add_argument('--m', type=int, default=8)
    ap.add_argument('--ks', type=int, default=256)
    ap.add_argument('--batch', type=int, default=10_000)
    ap.add_argument('--ef', type=int, default=64)
    ap.add_argument('--k', type=int, default=10)
    ap.add_argument('--M', type=int, default=16, help='HNSW max_connection')
    ap.add_argument('--efc', type=int, default=200)
    ap.add_argument('--dist', default='gaussian', choices=['gaussian', 'blobs'])
    ap.add_argument('--metric', default='euclidean')
    ap.add_argument('--mode', default='replicate', choices=['replicate', 'shard'])
    ap.add_argument('--build-threads', type=int, default=0)
    ap.add_argument('--no-cache', action='store_true')
    ap.add_argument('--cpu-sample', type=int, default=20000)
    ap.add_argument('--ref-sample', type=int, default=10000)
    ap.add_argument('--chunks', type=int, default=0, help='host-buffer pipeline depth inside annb_search (0 = auto)')
    ap.add_argument('--pool', type=int, default=4, help='distinct query batches cycled through the steps')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# synthetic data (seed 2)
# ---------------------------------------------------------------------------------------------------
def make_base(a, lo=0, hi=None):
    hi = a.n if hi is None else hi
    if a.dist == 'gaussian':
        # chunked so any [lo, hi) slice is reproducible without generating the whole matrix
        out = np.empty((hi - lo, a.dim), dtype=np.float32)
        CH = 100_000
        for c in range(lo // CH, (hi + CH - 1) // CH):
            rng = np.random.default_rng([2, c])
            blk = rng.standard_normal((CH, a.dim), dtype=np.float32)
            s, e = max(lo, c * CH), min(hi, (c + 1) * CH)
            out[s - lo:e - lo] = blk[s - c * CH:e - c * CH]
        return out
    rng = np.random.default_rng(123)   # examples/pq_benchmark.py:25-28 shape: 3 gaussian blobs
    centers = rng.uniform(-10, 10, (3, a.dim)).astype(np.float32)
    rng = np.random.default_rng([2, 77])
    lab = rng.integers(0, 3, a.n)
    x = centers[lab] + rng.standard_normal((a.n, a.dim), dtype=np.float32)
    return np.ascontiguousarray(x[lo:hi])


def make_queries(a, nb, rank=0):
    rng = np.random.default_rng([2, 1000 + rank])
    if a.dist == 'gaussian':
        return rng.standard_normal((nb, a.batch, a.dim), dtype=np.float32)
    crng = np.random.default_rng(123)
    centers = crng.uniform(-10, 10, (3, a.dim)).astype(np.float32)
    lab = rng.integers(0, 3, (nb, a.batch))
    return (centers[lab] + rng.standard_normal((nb, a.batch, a.dim), dtype=np.float32)).astype(np.float32)


def cfg_key(a, extra=''):
    s = f'{a.n}-{a.dim}-{a.m}-{a.ks}-{a.M}-{a.efc}-{a.dist}-{a.metric}-{extra}'
    return hashlib.md5(s.encode()).hexdigest()[:12]


def train_codebook(a, X10k):
    """PQ training is out of scope (an input to the path): sklearn KMeans like PQCodec.fit
    (pq.py:89-115) with the SURVEY 8d settings (random_state=0, n_init=1, max_iter=20)."""
    os.makedirs(CACHE, exist_ok=True)
    p = os.path.join(CACHE, f'codebook_{cfg_key(a)}.npy')
    if os.path.exists(p) and not a.no_cache:
        return np.load(p)
    from sklearn.cluster import KMeans
    ds = a.dim // a.m
    if a.metric == 'cosine':
        X10k = X10k / np.maximum(np.linalg.norm(X10k, axis=1, keepdims=True), 1e-12)
    cb = np.empty((a.m, a.ks, ds), dtype=np.float32)
    for m in range(a.m):
        km = KMeans(n_clusters=a.ks, max_iter=20, n_init=1, random_state=0).fit(X10k[:, m * ds:(m + 1) * ds])
        cb[m] = km.cluster_centers_
    np.save(p, cb)
    return cb


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu=0):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={self.gpu}', f'--query-gpu={self.Q}',
                                       '--format=csv,noheader,nounits', '-lms', '20'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.p:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        time.sleep(0.15)
        self.p.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 8 and r[4 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def ncu_traffic_bytes():
    """DRAM bytes per K3 launch from the committed ncu capture of this same command (profiles/)."""
    p = os.path.join(ROOT, 'profiles', 'r01_k3_walk_ncu.txt')
    try:
        tot = 0.0
        for line in open(p):
            f = line.split()
            if f and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                tot += float(f[1]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[f[2]]
        return tot or None
    except Exception:
        return None


def recall_at_k(pred, truth):
    return float(np.mean([len(set(p.tolist()) & set(t.tolist())) / len(t) for p, t in zip(pred, truth)]))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from annlite_b200.engine import Engine

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ncores = os.cpu_count() or 1
    shard = a.mode == 'shard' and world > 1

    # ---- index ---------------------------------------------------------------------------------
    cb = train_codebook(a, make_base(a, 0, 10_000)) if rank == 0 else None
    if world > 1:
        obj = [cb]
        dist.broadcast_object_list(obj, src=0)
        cb = obj[0]
    e = Engine(a.dim, a.m, a.ks, a.metric, device=local)
    e.set_codebook(cb)
    if a.chunks:
        e.set_option('chunks', a.chunks)
    os.makedirs(CACHE, exist_ok=True)
    t_build = 0.0
    if shard:
        lo, hi = rank * a.n // world, (rank + 1) * a.n // world
        path = os.path.join(CACHE, f'ours_{cfg_key(a, f"shard{rank}of{world}")}.hnsw')
        if os.path.exists(path) and not a.no_cache:
            e.load_index(path)
        else:
            X = make_base(a, lo, hi)
            e.init_graph(hi - lo, M=a.M, ef_construction=a.efc)
            t0 = time.time()
            e.add_items(X, np.arange(lo, hi, dtype=np.uint64), num_threads=a.build_threads or min(32, max(1, ncores // world)))
            t_build = time.time() - t0
            e.save_index(path)
            del X
    else:
        path = os.path.join(CACHE, f'ours_{cfg_key(a)}.hnsw')
        if rank == 0 and (a.no_cache or not os.path.exists(path)):
            X = make_base(a)
            e.init_graph(a.n, M=a.M, ef_construction=a.efc)
            t0 = time.time()
            e.add_items(X, np.arange(a.n, dtype=np.uint64), num_threads=a.build_threads)   # 0 = library default (<= 32, quota-aware)
            t_build = time.time() - t0
            e.save_index(path + '.tmp')
            os.replace(path + '.tmp', path)
            del X
        if world > 1:
            dist.barrier()
        if e.element_count == 0:
            e.load_index(path)

    # ---- queries --------------------------------------------------------------------------------
    nb = max(1, min(a.pool, a.steps + a.warmup))
    Qh = make_queries(a, nb, rank=0 if shard else rank)            # shard mode: same queries on every rank
    Qd = torch.from_numpy(Qh).cuda()
    B, k = a.batch, a.k
    out_l = torch.empty((B, k), dtype=torch.int64, device='cuda')
    out_d = torch.empty((B, k), dtype=torch.float32, device='cuda')
    if shard:
        g_l = torch.empty((world, B, k), dtype=torch.int64, device='cuda')
        g_d = torch.empty((world, B, k), dtype=torch.float32, device='cuda')
        m_l = torch.empty((B, k), dtype=torch.int64, device='cuda')
        m_d = torch.empty((B, k), dtype=torch.float32, device='cuda')
    stream = torch.cuda.ExternalStream(e.stream)
    norm = 2 if a.metric == 'cosine' else 0

    def step_dev(i):
        e.search(queries=Qd[i % nb], k=k, ef=a.ef, normalize=norm, out_labels=out_l, out_dists=out_d)
        if shard:
            dist.all_gather_into_tensor(g_l, out_l)
            dist.all_gather_into_tensor(g_d, out_d)
            torch.cuda.current_stream().synchronize()
            e.merge_topk(g_l, g_d, m_l, m_d)

    # pinned host buffers for the end-to-end leg
    Qp = torch.from_numpy(Qh).pin_memory()
    hl = torch.empty((B, k), dtype=torch.int64).pin_memory()
    hd = torch.empty((B, k), dtype=torch.float32).pin_memory()
    Qp_np, hl_np, hd_np = Qp.numpy(), hl.numpy().view(np.uint64), hd.numpy()

    def step_e2e(i):
        e.search(queries=Qp_np[i % nb], k=k, ef=a.ef, normalize=norm, out_labels=hl_np, out_dists=hd_np)
        if shard:   # host results -> device -> all-gather -> merge -> host
            out_l.copy_(hl, non_blocking=True)
            out_d.copy_(hd, non_blocking=True)
            dist.all_gather_into_tensor(g_l, out_l)
            dist.all_gather_into_tensor(g_d, out_d)
            torch.cuda.current_stream().synchronize()
            e.merge_topk(g_l, g_d, m_l, m_d)
            e.sync()
            hl.copy_(m_l)
            hd.copy_(m_d)

    def timed(fn, steps, warmup, drain=None):
        for i in range(warmup):
            fn(i)
        if drain:
            drain()
        e.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = e.launch_count
        t0 = time.perf_counter()
        ev0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        if drain:
            drain()
        e.sync()
        ev1.record(stream)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = max(ev0.elapsed_time(ev1), 0.0)
        # the two internal lanes finish on different streams: the CUDA-event span on lane 0 can end before
        # lane 1's last batch, so the step time is the larger of the event span and the host span that
        # encloses the final synchronisation of both lanes
        ms = max(ms, wall_ms) if drain else ms
        launches = e.launch_count - l0
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- streamed (submit / wait, two batches in flight) variants: the serving-loop shape ----
    out_l2 = [torch.empty((B, k), dtype=torch.int64, device='cuda') for _ in range(2)]
    out_d2 = [torch.empty((B, k), dtype=torch.float32, device='cuda') for _ in range(2)]
    hl2 = [torch.empty((B, k), dtype=torch.int64).pin_memory() for _ in range(2)]
    hd2 = [torch.empty((B, k), dtype=torch.float32).pin_memory() for _ in range(2)]
    hl2_np = [t.numpy().view(np.uint64) for t in hl2]
    hd2_np = [t.numpy() for t in hd2]
    pending = []

    def drain():
        while pending:
            e.search_wait(pending.pop(0))

    def step_dev_stream(i):
        if len(pending) == 2:
            e.search_wait(pending.pop(0))
        pending.append(e.search_submit(Qd[i % nb], out_l2[i & 1], out_d2[i & 1], k=k, ef=a.ef, normalize=norm))

    def step_e2e_stream(i):
        if len(pending) == 2:
            e.search_wait(pending.pop(0))
        pending.append(e.search_submit(Qp_np[i % nb], hl2_np[i & 1], hd2_np[i & 1], k=k, ef=a.ef, normalize=norm))

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.05)
    if shard:
        ms, launches = timed(step_dev, a.steps, a.warmup)
        ms_e2e, _ = timed(step_e2e, a.steps, max(3, a.warmup // 2))
        ms_sync, ms_e2e_sync = ms, ms_e2e
    else:
        ms, launches = timed(step_dev_stream, a.steps, a.warmup, drain)
        ms_e2e, _ = timed(step_e2e_stream, a.steps, max(3, a.warmup // 2), drain)
        ms_sync, _ = timed(step_dev, a.steps, 3)            # one blocking call per step, for reference
        ms_e2e_sync, _ = timed(step_e2e, a.steps, 3)
    ck = clocks.stop() if rank == 0 else None   # sampled across all timed regions
    # K3 duration per launch: CUDA events around the kernel on its own stream, blocking calls, untimed here
    kern_ms = 0.0
    for i in range(5):
        step_dev(i)
        kern_ms += e.last_kernel_ms()['search_ms']
    kern_ms *= a.steps / 5.0

    total_q = B * a.steps * (1 if shard else world)
    value = total_q / (ms / 1e3)
    e2e = total_q / (ms_e2e / 1e3)

    # ---- untimed: work counters (roofline), recall, CPU baseline ------------------------------------
    result = None
    if rank == 0 or shard:
        labels, dists, stats = e.search(queries=Qh[0], k=k, ef=a.ef, normalize=norm, with_stats=True)
    if rank == 0:
        hops, nbrs = stats[:, 0].astype(np.float64), stats[:, 1].astype(np.float64)
        code_row = a.m * (1 if a.ks <= 256 else 2)
        M0 = 2 * a.M
        # hops*(4+4*M0) + evals*(M*code_bytes) + evals/8 + k*12, plus the K3 read of the
        # materialised table (M*Ks*4) -- all per query; "evals" = neighbours listed (metric_distance_computations)
        alg_bytes_q = hops * (4 + 4 * M0) + nbrs * code_row + nbrs / 8 + k * 12 + a.m * a.ks * 4
        alg_bytes_launch = float(alg_bytes_q.sum())
        k3_ms = kern_ms / a.steps
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        achieved = alg_bytes_launch / (k3_ms / 1e3) / 1e9
        traffic = ncu_traffic_bytes() if (a.n == 1_000_000 and a.batch == 10_000) else None
        roof = {'bound': 'hbm', 'kernel': 'hnsw_walk_fast', 'achieved': round(achieved, 2), 'peak': peak,
                'peak_source': 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s',
                'unit': 'GB/s', 'frac': round(achieved / peak, 5), 'traffic': traffic,
                'traffic_source': 'profiles/r01_k3_walk_ncu.txt (ncu --set full of this command): dram__bytes_read.sum + dram__bytes_write.sum per launch',
                'alg_bytes_per_launch': round(alg_bytes_launch),
                'ms_per_launch': round(k3_ms, 4), 'alg_bytes_per_query': round(float(alg_bytes_q.mean()), 1),
                'hops_per_query': round(float(hops.mean()), 2), 'nbrs_per_query': round(float(nbrs.mean()), 1),
                'note': 'latency-bound pointer chase; moved bytes/query = hops*record(384B)+table(8KB)'}
        # recall vs exhaustive ADC on the same codes (GPU K2) and vs true L2 on a sample
        sample = min(1000, B)
        if not shard:
            g = e.get_graph()
            n = g['cur_element_count']
            codes = g['data_level0'].reshape(n, -1)[:, g['offset_data']:g['label_offset']].copy()
            lab = np.ascontiguousarray(g['data_level0'].reshape(n, -1)[:, g['label_offset']:g['label_offset'] + 8]).view(np.uint64).ravel()
            e.set_codes(codes.view(np.uint8 if a.ks <= 256 else np.uint16).reshape(n, a.m))
            tbl = e.adc_table(Qh[0][:sample], normalize=1 if a.metric == 'cosine' else 0) if a.metric == 'euclidean' else None
            rec_adc = None
            if tbl is not None:
                gt_i, _ = e.scan_topk(tables=tbl, k=k)
                rec_adc = recall_at_k(labels[:sample], lab[gt_i])
            Xd = torch.from_numpy(make_base(a)).cuda()
            qd = Qd[0][:sample]
            d2 = (qd * qd).sum(1, keepdim=True) - 2 * qd @ Xd.T + (Xd * Xd).sum(1)[None]
            gt = d2.topk(k, dim=1, largest=False).indices.cpu().numpy()
            rec_l2 = recall_at_k(labels[:sample].astype(np.int64), gt)
            del Xd, d2
        else:
            rec_adc = rec_l2 = None
        # CPU baseline: C oracle port over the same graph, all host threads, bounded sample
        cpu = cpu_port_baseline(a, e, Qh, cb, ncores) if world == 1 else None   # reported at N=1 only
        result = {
            'metric': 'queries/sec (PQ-HNSW search, 1M x 128d, M=8, ef=64, k=10)', 'value': round(value, 1),
            'unit': 'queries/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': round(ms / a.steps, 4), 'higher_is_better': True, 'scaling': 'strong' if shard else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'configs[1]: {a.n}x{a.dim} fp32 {a.dist}, PQ M={a.m} Ks={a.ks}, HNSW M={a.M} '
                                   f'efc={a.efc}, ef={a.ef}, k={a.k}, batch={a.batch} queries/step/GPU',
                       'parallelism': ('shard' if shard else 'replicate') + str(world),
                       'l2_policy': 'index (384 MB walk records) exceeds the 126 MB L2; query batches rotate '
                                    f'through a pool of {nb}',
                       'metric_space': a.metric, 'index_build_s': round(t_build, 1), 'host_cores': ncores,
                       'api': ('streamed: annb_search_submit/wait with two batches in flight (value and e2e); '
                               'blocking_call_value = one annb_search call at a time') if not shard else 'blocking annb_search'},
            'e2e': {'value': round(e2e, 1), 'unit': 'queries/s', 'h2d_bytes_per_step': B * a.dim * 4,
                    'd2h_bytes_per_step': B * k * 12 + B * 4, 'ms_per_step': round(ms_e2e / a.steps, 4),
                    'api': 'annb_search_submit/wait, 2 batches in flight' if not shard else 'annb_search + all-gather + merge',
                    'blocking_call_value': round(total_q / (ms_e2e_sync / 1e3), 1)},
            'blocking_call_value': round(total_q / (ms_sync / 1e3), 1),
            'gpu_launches': int(launches), 'clocks': ck, 'roofline': roof, 'cpu_baseline': cpu,
            'recall_at_k': {'vs_exhaustive_adc': rec_adc, 'vs_true_l2': rec_l2, 'sample': sample},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def cpu_port_baseline(a, e, Qh, cb, ncores):
    """oracle port (C restatement, oracle/pq_oracle.c) on the host cores: tables + walk, threads over
    query slices (ctypes releases the GIL).  Bounded sample of the same query batch."""
    import oracle as O
    g = O.Graph.from_state(e.get_graph(), a.m, a.ks)
    S = min(a.cpu_sample, Qh.shape[0] * Qh.shape[1])
    q = Qh.reshape(-1, a.dim)[:S]
    T = min(ncores, 64)
    parts = np.array_split(np.arange(S), T)

    def work(idx):
        t = O.adc_table(q[idx], cb, a.metric)
        O.hnsw_search(g, t, a.k, a.ef)

    work(parts[0][:64])  # warm
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(p,)) for p in parts if len(p)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    return {'value': round(S / dt, 1), 'unit': 'queries/s', 'cores': T, 'kind': 'port',
            'sample': f'{S} queries of the same batch (tables + walk), {T} threads, {dt:.2f}s'}


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------
def run_reference(a):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import ref_driver as R
    if not R.available():
        print(json.dumps({'impl': 'reference', 'unavailable': 'oracle/_ref not built (needs /root/reference at build time)'}))
        return
    ncores = os.cpu_count() or 1
    world = int(os.environ.get('WORLD_SIZE', 1))
    cb = train_codebook(a, make_base(a, 0, 10_000))
    codec = R.RefCodec(cb, a.metric)
    X = make_base(a)
    idx = R.RefHnswIndex(codec, a.metric, capacity=a.n, ef_construction=a.efc, ef_search=a.ef, max_connection=a.M)
    t0 = time.time()
    idx.add_with_ids(X, np.arange(a.n), num_threads=ncores, batch=5000)
    t_build = time.time() - t0
    nb = max(1, min(a.pool, a.steps + a.warmup))
    Qh = make_queries(a, nb, rank=0)
    # bounded sample per step so that the whole --steps/--warmup run stays within a couple of minutes
    S = min(a.ref_sample, a.batch, max(256, 1_500_000 // max(1, a.steps + a.warmup)))

    def step(i):
        q = Qh[i % nb][:S]
        tables = codec.get_dist_mat(idx._pre(q))           # pq_bind (single-threaded Cython)
        return idx.knn_query(q, a.k, num_threads=ncores, tables=tables)   # hnsw_bind, all threads

    for i in range(a.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        labels, dists = step(a.warmup + i)
    dt = time.perf_counter() - t0
    qps = S * a.steps / dt
    # walk-only rate (tables prebuilt) for context
    q = Qh[0][:S]
    tables = codec.get_dist_mat(idx._pre(q))
    t1 = time.perf_counter()
    idx.knn_query(q, a.k, num_threads=ncores, tables=tables)
    walk_qps = S / (time.perf_counter() - t1)
    # as-shipped semantics: one query per call (AnnLite.search loop), bounded
    t2 = time.perf_counter()
    for j in range(300):
        idx.search(Qh[0][j], limit=a.k)
    loop_qps = 300 / (time.perf_counter() - t2)
    res = {
        'impl': 'reference', 'metric': 'queries/sec (PQ-HNSW search, 1M x 128d, M=8, ef=64, k=10)',
        'value': round(qps, 1), 'unit': 'queries/s', 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
        'ms_per_step': round(dt / a.steps * 1e3, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'configs[1]: {a.n}x{a.dim} fp32 {a.dist}, PQ M={a.m} Ks={a.ks}, HNSW M={a.M} '
                               f'efc={a.efc}, ef={a.ef}, k={a.k}; each step = {S} queries (bounded sample of the '
                               f'{a.batch}-query batch)', 'index_build_s': round(t_build, 1), 'host_cores': ncores,
                   'walk_only_qps': round(walk_qps, 1), 'one_query_per_call_qps': round(loop_qps, 1)},
        'cpu_baseline': {'value': round(qps, 1), 'unit': 'queries/s', 'cores': ncores, 'kind': 'reference',
                         'sample': f'{S} queries/step: pq_bind.batch_precompute_adc_table (1 thread) + '
                                   f'hnsw_bind.Index.knn_query ({ncores} threads)'},
        'e2e': {'value': round(qps, 1), 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(res))


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
