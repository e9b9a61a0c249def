// adc_scan.cu -- K2: exhaustive ADC scan, plain (dist_pqcodes_to_codebooks) and fused with top-k
// (PQIndex.search), plus the (dist, id) k-way merge used for chunk partials and shard results.
//
// d[n] = sum_m T[m, code[n,m]], m sequential from 0.f     bindings/pq_bindings.pyx:30-47, :52-80
// top-k: k smallest ascending, ties by row index          pq_index.py:29-56 + annlite/math.py:94-120
//        (numpy leaves tie order unspecified; (dist, index) is the rule the oracle fixes too)
//
// Mapping (scan_topk): a CTA stages QT query tables in shared memory (one per warp) and streams a
// chunk of the code matrix; each lane scores one code row per step (coalesced code loads, L1
// shared by the CTA's warps), and the warp keeps its query's running top-k in a register-
// resident sorted list with a threshold test, so inserts are rare after warm-up.  Partials
// (B, chunks, k) are merged by merge_topk_kernel.  Bound: shared-memory gather rate for N*M
// small (C1: 80 KB of codes, L2/L1 resident), HBM reads of N*M bytes per query tile otherwise.
#include <math_constants.h>

#include "annb_internal.h"
#include "warp_list.cuh"

namespace {

template <typename code_t>
__device__ __forceinline__ float adc_lookup_row(const float *__restrict__ T, const code_t *__restrict__ row, int M,
                                                int Ks) {
  float r = 0.f;
  for (int m = 0; m < M; m++) r = __fadd_rn(r, T[m * Ks + (int)row[m]]);
  return r;
}

// specialisation for 8 one-byte codes (M = 8): one 64-bit load per row
__device__ __forceinline__ float adc_lookup_u8x8(const float *__restrict__ T, uint2 c, int Ks) {
  float r = 0.f;
  r = __fadd_rn(r, T[0 * Ks + (c.x & 0xff)]);
  r = __fadd_rn(r, T[1 * Ks + ((c.x >> 8) & 0xff)]);
  r = __fadd_rn(r, T[2 * Ks + ((c.x >> 16) & 0xff)]);
  r = __fadd_rn(r, T[3 * Ks + (c.x >> 24)]);
  r = __fadd_rn(r, T[4 * Ks + (c.y & 0xff)]);
  r = __fadd_rn(r, T[5 * Ks + ((c.y >> 8) & 0xff)]);
  r = __fadd_rn(r, T[6 * Ks + ((c.y >> 16) & 0xff)]);
  r = __fadd_rn(r, T[7 * Ks + (c.y >> 24)]);
  return r;
}

template <typename code_t>
__global__ void scan_kernel(const float *__restrict__ table, const code_t *__restrict__ codes, float *__restrict__ out,
                            int64_t N, int M, int Ks, int table_in_smem) {
  extern __shared__ float sT[];
  const float *T = table;
  if (table_in_smem) {
    for (int i = threadIdx.x; i < M * Ks; i += blockDim.x) sT[i] = table[i];
    __syncthreads();
    T = sT;
  }
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
    float d;
    if (sizeof(code_t) == 1 && M == 8) {
      uint2 c = *reinterpret_cast<const uint2 *>(codes + n * 8);
      d = adc_lookup_u8x8(T, c, Ks);
    } else {
      d = adc_lookup_row<code_t>(T, codes + n * M, M, Ks);
    }
    out[n] = d;
  }
}

// partial top-k of one (query, chunk): results to part_d/part_i [(b*chunks + chunk)*k ...]
template <int EPL, typename code_t>
__global__ void __launch_bounds__(256)
scan_topk_kernel(const float *__restrict__ tables, const code_t *__restrict__ codes, int64_t B, int64_t N, int M, int Ks,
                 int k, int chunks, int64_t rows_per_chunk, int table_in_smem, float *__restrict__ part_d,
                 uint32_t *__restrict__ part_i) {
  extern __shared__ float sT[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = blockDim.x >> 5;
  const int64_t b = (int64_t)blockIdx.y * W + warp;
  const int chunk = blockIdx.x;
  const int TS = M * Ks;
  const float *T = nullptr;
  if (b < B) {
    T = tables + b * TS;
    if (table_in_smem) {
      float *dst = sT + warp * TS;
      for (int i = lane; i < TS; i += 32) dst[i] = T[i];
      T = dst;
    }
  }
  __syncwarp();
  if (b >= B) return;
  WarpList<EPL> L;
  L.clear();
  float worst = CUDART_INF_F;
  const int64_t r0 = (int64_t)chunk * rows_per_chunk;
  const int64_t r1 = min(N, r0 + rows_per_chunk);
  for (int64_t base = r0; base < r1; base += 32) {
    const int64_t n = base + lane;
    float d = CUDART_INF_F;
    if (n < r1) {
      if (sizeof(code_t) == 1 && M == 8) {
        uint2 c = *reinterpret_cast<const uint2 *>(codes + n * 8);
        d = adc_lookup_u8x8(T, c, Ks);
      } else {
        d = adc_lookup_row<code_t>(T, codes + n * M, M, Ks);
      }
    }
    unsigned mask = __ballot_sync(FULL_MASK, d < worst);
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const float dj = __shfl_sync(FULL_MASK, d, j);
      if (dj < worst) {
        L.insert(dj, (uint32_t)(base + j - r0), k);
        worst = L.key_at(k - 1);
      }
    }
  }
  // write the k best (positions 0..k-1), row index made global
#pragma unroll
  for (int e = 0; e < EPL; e++) {
    const int pos = e * 32 + lane;
    if (pos < k) {
      const int64_t o = (b * chunks + chunk) * k + pos;
      part_d[o] = L.k[e];
      part_i[o] = L.v[e] == LIST_EMPTY_VAL ? LIST_EMPTY_VAL : (uint32_t)(L.v[e] + r0);
    }
  }
}

// ---- query-tiled scan (round 2): lanes = QUERIES, not rows --------------------------------------------------
// scan_topk_kernel gives every warp its own table and lets its 32 lanes look up 32 DIFFERENT codes per step: 32
// random words of a 256-entry row land in the 32 banks ~3.5 deep, and the measured rate is 15 % of the shared-memory
// peak (bench.py k2_exhaustive_scan, 1M rows).  Here a CTA stages the tables of QT = 16 queries INTERLEAVED --
// sT[(m*Ks + c)*16 + q] -- and a half-warp scores ONE row for those 16 queries: its 16 lanes read 16 consecutive
// words (the code is the same for all of them), so a warp instruction touches two 64-byte runs: conflict-free or
// 2-way.  The row's 8 code bytes are loaded once per 16 queries.  Every lane keeps the running top-KK of ITS
// (query, row-parity) pair in registers (static indexing, bubble insertion -- rare once the threshold has settled);
// the two halves are merged by shuffles, and the per-warp lists (G = chunks x warps per query) go to
// merge_topk_kernel.  Same arithmetic (m sequential, fp32), same (dist, row) order: bit-identical results.
constexpr int QT = 16;
constexpr int TILED_WARPS = 16;

template <int KK>
__device__ __forceinline__ void lane_list_insert(float (&Kd)[KK], uint32_t (&Ki)[KK], float d, uint32_t id) {
  if (d < Kd[KK - 1] || (d == Kd[KK - 1] && id < Ki[KK - 1])) {
    Kd[KK - 1] = d;
    Ki[KK - 1] = id;
#pragma unroll
    for (int i = KK - 1; i > 0; i--) {
      const bool sw = Kd[i] < Kd[i - 1] || (Kd[i] == Kd[i - 1] && Ki[i] < Ki[i - 1]);
      const float td = Kd[i];
      const uint32_t ti = Ki[i];
      Kd[i] = sw ? Kd[i - 1] : td;
      Ki[i] = sw ? Ki[i - 1] : ti;
      Kd[i - 1] = sw ? td : Kd[i - 1];
      Ki[i - 1] = sw ? ti : Ki[i - 1];
    }
  }
}

template <int KK>
__global__ void __launch_bounds__(TILED_WARPS * 32, 1)
scan_topk_tiled_kernel(const float *__restrict__ tables, const uint8_t *__restrict__ codes, int64_t B, int64_t N, int Ks, int k,
                       int chunks, int64_t rows_per_chunk, float *__restrict__ part_d, uint32_t *__restrict__ part_i) {
  constexpr int M = 8;
  extern __shared__ float sT[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, q = lane & 15;
  const int64_t b0 = (int64_t)blockIdx.y * QT;
  const int chunk = blockIdx.x;
  const int TS = M * Ks;
  // stage the tile: thread -> (table entry i, query q), q fastest => conflict-free shared-memory stores
  for (int t = threadIdx.x; t < TS * QT; t += blockDim.x) {
    const int qq = t & (QT - 1), i = t >> 4;
    sT[t] = (b0 + qq < B) ? __ldg(tables + (b0 + qq) * TS + i) : 0.f;
  }
  __syncthreads();
  float Kd[KK];
  uint32_t Ki[KK];
#pragma unroll
  for (int i = 0; i < KK; i++) {
    Kd[i] = CUDART_INF_F;
    Ki[i] = LIST_EMPTY_VAL;
  }
  const int64_t r0 = (int64_t)chunk * rows_per_chunk;
  const int64_t r1 = min(N, r0 + rows_per_chunk);
  const float *Tq = sT + q;
  auto score = [&](const uint2 c) -> float {  // m sequential from 0.f, every add rounded (pq_bindings.pyx:30-47)
    float d = 0.f;
    d = __fadd_rn(d, Tq[(0 * Ks + (int)(c.x & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(1 * Ks + (int)((c.x >> 8) & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(2 * Ks + (int)((c.x >> 16) & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(3 * Ks + (int)(c.x >> 24)) * QT]);
    d = __fadd_rn(d, Tq[(4 * Ks + (int)(c.y & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(5 * Ks + (int)((c.y >> 8) & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(6 * Ks + (int)((c.y >> 16) & 0xff)) * QT]);
    d = __fadd_rn(d, Tq[(7 * Ks + (int)(c.y >> 24)) * QT]);
    return d;
  };
  // two rows per lane and iteration: two independent lookup chains in flight (16 warps per SM is all the tile leaves room for)
  constexpr int64_t STEP = 2 * TILED_WARPS;
  int64_t r = r0 + 2 * warp + half;
  for (; r + STEP < r1; r += 2 * STEP) {
    const uint2 ca = __ldg(reinterpret_cast<const uint2 *>(codes + r * 8));
    const uint2 cb = __ldg(reinterpret_cast<const uint2 *>(codes + (r + STEP) * 8));
    const float da = score(ca), db = score(cb);
    lane_list_insert<KK>(Kd, Ki, da, (uint32_t)r);
    lane_list_insert<KK>(Kd, Ki, db, (uint32_t)(r + STEP));
  }
  if (r < r1) {
    const uint2 ca = __ldg(reinterpret_cast<const uint2 *>(codes + r * 8));
    lane_list_insert<KK>(Kd, Ki, score(ca), (uint32_t)r);
  }
  // the odd-row half hands its list to the even-row half of the same query
#pragma unroll
  for (int i = 0; i < KK; i++) {
    const float od = __shfl_down_sync(FULL_MASK, Kd[i], 16);
    const uint32_t oi = __shfl_down_sync(FULL_MASK, Ki[i], 16);
    if (half == 0 && oi != LIST_EMPTY_VAL) lane_list_insert<KK>(Kd, Ki, od, oi);
  }
  if (half == 0 && b0 + q < B) {
    const int64_t G = (int64_t)chunks * TILED_WARPS;
    const int64_t o = ((b0 + q) * G + (int64_t)chunk * TILED_WARPS + warp) * k;
#pragma unroll
    for (int i = 0; i < KK; i++)
      if (i < k) {
        part_d[o + i] = Kd[i];
        part_i[o + i] = Ki[i];
      }
  }
}

// ---- the same, on a diet (ncu of the first version: 73 warp instructions per two-row step, 4.5 ms for 1000 queries x
// 1M rows, and 3 ms more in the merge of the 16 x chunks per-warp lists): Ks = 256 at compile time (lookup = byte
// extract + shift + LDS with an immediate row offset), the threshold test of the streaming loop compares distances only
// (rows arrive in ascending order, so an equal distance never displaces), and the CTA reduces its 16 per-warp lists
// to ONE list per (query, chunk) through shared memory before anything is written.
template <int KK>
__device__ __forceinline__ void lane_list_insert_ascending_rows(float (&Kd)[KK], uint32_t (&Ki)[KK], float d, uint32_t id) {
  if (d < Kd[KK - 1]) {
    Kd[KK - 1] = d;
    Ki[KK - 1] = id;
#pragma unroll
    for (int i = KK - 1; i > 0; i--) {
      const bool sw = Kd[i] < Kd[i - 1];
      const float td = Kd[i];
      const uint32_t ti = Ki[i];
      Kd[i] = sw ? Kd[i - 1] : td;
      Ki[i] = sw ? Ki[i - 1] : ti;
      Kd[i - 1] = sw ? td : Kd[i - 1];
      Ki[i - 1] = sw ? ti : Ki[i - 1];
    }
  }
}

template <int KK>
__global__ void __launch_bounds__(TILED_WARPS * 32, 1)
scan_topk_tiled2_kernel(const float *__restrict__ tables, const uint8_t *__restrict__ codes, int64_t B, int64_t N, int k, int chunks,
                        int64_t rows_per_chunk, float *__restrict__ part_d, uint32_t *__restrict__ part_i) {
  constexpr int M = 8, KS = 256;
  extern __shared__ float sT[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, q = lane & 15;
  const int64_t b0 = (int64_t)blockIdx.y * QT;
  const int chunk = blockIdx.x;
  constexpr int TS = M * KS;
  for (int t = threadIdx.x; t < TS * QT; t += blockDim.x) {
    const int qq = t & (QT - 1), i = t >> 4;
    sT[t] = (b0 + qq < B) ? __ldg(tables + (b0 + qq) * TS + i) : 0.f;
  }
  __syncthreads();
  float Kd[KK];
  uint32_t Ki[KK];
#pragma unroll
  for (int i = 0; i < KK; i++) {
    Kd[i] = CUDART_INF_F;
    Ki[i] = LIST_EMPTY_VAL;
  }
  const int64_t r0 = (int64_t)chunk * rows_per_chunk;
  const int64_t r1 = min(N, r0 + rows_per_chunk);
  const float *Tq = sT + q;
  auto score = [&](const uint2 c) -> float {  // m sequential from 0.f, every add rounded (pq_bindings.pyx:30-47)
    float d = 0.f;
    d = __fadd_rn(d, Tq[((c.x & 0xffu) + 0 * KS) * QT]);
    d = __fadd_rn(d, Tq[(((c.x >> 8) & 0xffu) + 1 * KS) * QT]);
    d = __fadd_rn(d, Tq[(((c.x >> 16) & 0xffu) + 2 * KS) * QT]);
    d = __fadd_rn(d, Tq[((c.x >> 24) + 3 * KS) * QT]);
    d = __fadd_rn(d, Tq[((c.y & 0xffu) + 4 * KS) * QT]);
    d = __fadd_rn(d, Tq[(((c.y >> 8) & 0xffu) + 5 * KS) * QT]);
    d = __fadd_rn(d, Tq[(((c.y >> 16) & 0xffu) + 6 * KS) * QT]);
    d = __fadd_rn(d, Tq[((c.y >> 24) + 7 * KS) * QT]);
    return d;
  };
  constexpr int64_t STEP = 2 * TILED_WARPS;
  int64_t r = r0 + 2 * warp + half;
  for (; r + STEP < r1; r += 2 * STEP) {
    const uint2 ca = __ldg(reinterpret_cast<const uint2 *>(codes + r * 8));
    const uint2 cb = __ldg(reinterpret_cast<const uint2 *>(codes + (r + STEP) * 8));
    const float da = score(ca), db = score(cb);
    lane_list_insert_ascending_rows<KK>(Kd, Ki, da, (uint32_t)r);
    lane_list_insert_ascending_rows<KK>(Kd, Ki, db, (uint32_t)(r + STEP));
  }
  if (r < r1) {
    const uint2 ca = __ldg(reinterpret_cast<const uint2 *>(codes + r * 8));
    lane_list_insert_ascending_rows<KK>(Kd, Ki, score(ca), (uint32_t)r);
  }
  // the odd-row half hands its list to the even-row half of the same query: (dist, row) order from here on
#pragma unroll
  for (int i = 0; i < KK; i++) {
    const float od = __shfl_down_sync(FULL_MASK, Kd[i], 16);
    const uint32_t oi = __shfl_down_sync(FULL_MASK, Ki[i], 16);
    if (half == 0 && oi != LIST_EMPTY_VAL) lane_list_insert<KK>(Kd, Ki, od, oi);
  }
  // ---- CTA reduction: the 16 per-warp lists of a query -> one list, through the (no longer needed) table area ----
  __syncthreads();  // every warp is done with the tables
  float *sd = sT;
  uint32_t *si = reinterpret_cast<uint32_t *>(sT + QT * TILED_WARPS * KK);
  if (half == 0) {
#pragma unroll
    for (int i = 0; i < KK; i++) {
      sd[(q * TILED_WARPS + warp) * KK + i] = Kd[i];
      si[(q * TILED_WARPS + warp) * KK + i] = Ki[i];
    }
  }
  __syncthreads();
  // warp w finishes query w: lane l < 16 holds the list warp l produced, the other lanes hold empty lists
#pragma unroll
  for (int i = 0; i < KK; i++) {
    Kd[i] = half == 0 ? sd[(warp * TILED_WARPS + q) * KK + i] : CUDART_INF_F;
    Ki[i] = half == 0 ? si[(warp * TILED_WARPS + q) * KK + i] : LIST_EMPTY_VAL;
  }
  const int64_t b = b0 + warp;
  const int64_t o = (b * chunks + chunk) * k;
  for (int j = 0; j < k; j++) {
    float md = Kd[0];
    uint32_t mi = Ki[0];
#pragma unroll
    for (int s2 = 16; s2 > 0; s2 >>= 1) {
      const float od = __shfl_xor_sync(FULL_MASK, md, s2);
      const uint32_t oi = __shfl_xor_sync(FULL_MASK, mi, s2);
      if (od < md || (od == md && oi < mi)) {
        md = od;
        mi = oi;
      }
    }
    if (lane == 0 && b < B) {
      part_d[o + j] = md;
      part_i[o + j] = mi;
    }
    if (Ki[0] == mi && mi != LIST_EMPTY_VAL) {  // row ids are unique across the lists: exactly one lane pops its head
#pragma unroll
      for (int i = 0; i + 1 < KK; i++) {
        Kd[i] = Kd[i + 1];
        Ki[i] = Ki[i + 1];
      }
      Kd[KK - 1] = CUDART_INF_F;
      Ki[KK - 1] = LIST_EMPTY_VAL;
    }
  }
}

// Merge G sorted (ascending) lists of k (dist, id) per query into the global k best, ordered
// by (dist, id).  One warp per query; lists are tiny (G*k entries).  Used for chunk partials
// (ids = u32 row index -> i64) and for shard results (ids = u64 labels).
// CellContainer.ivf_search merge: annlite/container.py:130-138 (hstack -> argsort[:limit]).
template <typename in_id_t, typename out_id_t, int EPL>
__global__ void merge_topk_kernel(const float *__restrict__ d_in, const in_id_t *__restrict__ i_in, int G, int64_t B, int k,
                                  int64_t g_stride, int64_t b_stride, float *__restrict__ d_out,
                                  out_id_t *__restrict__ i_out, in_id_t empty_in, out_id_t empty_out) {
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  // Rank-based merge: entry x gets rank = #{y : (d_y, id_y) < (d_x, id_x)}; ranks < k are written.
  // G*k is small (<= a few thousand), O((G*k)^2 / 32) per query is negligible next to the scan / walk.
  const int total = G * k;
  for (int x = lane; x < total; x += 32) {
    const int gx = x / k, px = x - gx * k;
    const float dx = d_in[gx * g_stride + b * b_stride + px];
    const in_id_t ix = i_in[gx * g_stride + b * b_stride + px];
    if (ix == empty_in) continue;
    int rank = 0;
    for (int y = 0; y < total; y++) {
      const int gy = y / k, py = y - gy * k;
      const float dy = d_in[gy * g_stride + b * b_stride + py];
      const in_id_t iy = i_in[gy * g_stride + b * b_stride + py];
      if (iy == empty_in) continue;
      rank += (dy < dx) || (dy == dx && (iy < ix || (iy == ix && y < x)));
    }
    if (rank < k) {
      d_out[b * k + rank] = dx;
      i_out[b * k + rank] = (out_id_t)ix;
    }
  }
  // fill the tail when fewer than k valid entries exist
  int valid = 0;
  for (int x = lane; x < total; x += 32) {
    const int gx = x / k, px = x - gx * k;
    valid += i_in[gx * g_stride + b * b_stride + px] != empty_in;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(FULL_MASK, valid, o);
  for (int p = valid + lane; p < k; p += 32) {
    d_out[b * k + p] = CUDART_INF_F;
    i_out[b * k + p] = empty_out;
  }
}

}  // namespace

static size_t smem_optin_limit(int device) {
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  return (size_t)v;
}

int launch_scan(annb_index *h, const float *d_table, float *d_out) {
  const int64_t N = h->n_codes;
  if (N == 0) return ANNB_OK;
  const size_t tbytes = (size_t)h->M * h->Ks * sizeof(float);
  const int in_smem = tbytes <= 96 * 1024;
  int blocks = (int)std::min<int64_t>((N + 255) / 256, (int64_t)h->sm_count * 8);
  if (h->code_bytes == 1) {
    if (in_smem && tbytes > 48 * 1024)
      ANNB_CUDA(cudaFuncSetAttribute(scan_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbytes));
    scan_kernel<uint8_t><<<blocks, 256, in_smem ? tbytes : 0, h->stream>>>(d_table, (const uint8_t *)h->d_codes, d_out, N,
                                                                         h->M, h->Ks, in_smem);
  } else {
    if (in_smem && tbytes > 48 * 1024)
      ANNB_CUDA(cudaFuncSetAttribute(scan_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tbytes));
    scan_kernel<uint16_t><<<blocks, 256, in_smem ? tbytes : 0, h->stream>>>(d_table, (const uint16_t *)h->d_codes, d_out,
                                                                          N, h->M, h->Ks, in_smem);
  }
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

template <int EPL, typename code_t>
static int run_scan_topk(annb_index *h, const float *d_tables, int64_t B, int k, int chunks, int64_t rows_per_chunk,
                         int W, int in_smem, size_t smem, float *part_d, uint32_t *part_i) {
  auto kern = scan_topk_kernel<EPL, code_t>;
  if (smem > 48 * 1024) ANNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t done = 0;
  while (done < B) {
    int64_t nb = std::min<int64_t>(B - done, (int64_t)65535 * W);
    dim3 grid((unsigned)chunks, (unsigned)((nb + W - 1) / W));
    kern<<<grid, W * 32, smem, h->stream>>>(d_tables + done * h->M * h->Ks, (const code_t *)h->d_codes, nb, h->n_codes,
                                            h->M, h->Ks, k, chunks, rows_per_chunk, in_smem,
                                            part_d + done * chunks * k, part_i + done * chunks * k);
    h->launches++;
    ANNB_CUDA(cudaGetLastError());
    done += nb;
  }
  return ANNB_OK;
}

// the query-tiled scan: M = 8 one-byte codes, a 16-query interleaved tile that fits shared memory, k <= 16 (per-lane
// register lists), enough queries to fill tiles and enough rows to amortise staging a 128 KB tile
template <int KK>
static int run_scan_topk_tiled(annb_index *h, const float *d_tables, int64_t B, int k, int64_t *d_ids, float *d_dists) {
  const int64_t N = h->n_codes;
  const size_t smem = (size_t)h->M * h->Ks * QT * sizeof(float);
  const bool v2 = h->Ks == 256 && h->opt_scan_kernel != 3;   // the diet version: Ks = 256, one list per (query, chunk)
  auto kern = scan_topk_tiled_kernel<KK>;
  auto kern2 = scan_topk_tiled2_kernel<KK>;
  if (v2) ANNB_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else ANNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (B + QT - 1) / QT;
  // one CTA per SM is resident (the tile): ~2 waves of CTAs, at most 16 chunks (the merge sees chunks x 16 lists)
  int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(16, (2LL * h->sm_count + tiles - 1) / tiles));
  chunks = (int)std::min<int64_t>(chunks, std::max<int64_t>(1, N / 4096));
  int64_t rows_per_chunk = (N + chunks - 1) / chunks;
  rows_per_chunk = (rows_per_chunk + 31) / 32 * 32;
  chunks = (int)std::max<int64_t>(1, (N + rows_per_chunk - 1) / rows_per_chunk);
  const int G = v2 ? chunks : chunks * TILED_WARPS;
  float *part_d;
  uint32_t *part_i;
  int rc;
  if ((rc = annb_scratch(h, 16, (size_t)B * G * k * sizeof(float), (void **)&part_d))) return rc;
  if ((rc = annb_scratch(h, 17, (size_t)B * G * k * sizeof(uint32_t), (void **)&part_i))) return rc;
  int64_t done = 0;
  while (done < B) {
    const int64_t nb = std::min<int64_t>(B - done, (int64_t)65535 * QT);
    dim3 grid((unsigned)chunks, (unsigned)((nb + QT - 1) / QT));
    if (v2)
      kern2<<<grid, TILED_WARPS * 32, smem, h->stream>>>(d_tables + done * h->M * h->Ks, (const uint8_t *)h->d_codes, nb, N, k, chunks,
                                                         rows_per_chunk, part_d + done * G * k, part_i + done * G * k);
    else
      kern<<<grid, TILED_WARPS * 32, smem, h->stream>>>(d_tables + done * h->M * h->Ks, (const uint8_t *)h->d_codes, nb, N, h->Ks, k, chunks,
                                                        rows_per_chunk, part_d + done * G * k, part_i + done * G * k);
    h->launches++;
    ANNB_CUDA(cudaGetLastError());
    done += nb;
  }
  const int warps = 4;
  merge_topk_kernel<uint32_t, int64_t, 1><<<(unsigned)((B + warps - 1) / warps), warps * 32, 0, h->stream>>>(
      part_d, part_i, G, B, k, (int64_t)k, (int64_t)G * k, d_dists, d_ids, LIST_EMPTY_VAL, (int64_t)-1);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

int launch_scan_topk(annb_index *h, const float *d_tables, int64_t B, int k, int64_t *d_ids, float *d_dists) {
  if (B == 0) return ANNB_OK;
  const int64_t N = h->n_codes;
  if (k > ANNB_MAX_EF) ANNB_FAIL(ANNB_ELIMIT, "k=%d exceeds the limit %d of the register-resident top-k list", k, ANNB_MAX_EF);
  if (N >= (int64_t)0xffffffffll) ANNB_FAIL(ANNB_ELIMIT, "scan supports < 2^32-1 rows per index (shard larger sets)");
  const size_t tbytes = (size_t)h->M * h->Ks * sizeof(float);
  const size_t lim = smem_optin_limit(h->device);
  if (h->opt_scan_kernel != 1 && h->M == 8 && h->code_bytes == 1 && k <= 16 && tbytes * QT + 1024 <= lim &&
      ((B >= 64 && N >= 32768) || h->opt_scan_kernel >= 2)) {
    if (k <= 1) return run_scan_topk_tiled<1>(h, d_tables, B, k, d_ids, d_dists);
    if (k <= 10) return run_scan_topk_tiled<10>(h, d_tables, B, k, d_ids, d_dists);
    return run_scan_topk_tiled<16>(h, d_tables, B, k, d_ids, d_dists);
  }
  int W = 8;
  int in_smem = 1;
  while (W > 1 && W * tbytes > lim - 1024) W >>= 1;
  if (W * tbytes > lim - 1024) in_smem = 0, W = 8;
  const size_t smem = in_smem ? W * tbytes : 0;
  // enough (tile, chunk) CTAs to fill the GPU for ~2 waves; chunks of >= 1024 rows
  const int64_t tiles = (B + W - 1) / W;
  int64_t want = std::max<int64_t>(1, (2LL * h->sm_count * 2 + tiles - 1) / tiles);
  int64_t maxchunks = std::max<int64_t>(1, N / 1024);
  int chunks = (int)std::min<int64_t>(std::min<int64_t>(want, maxchunks), 128);
  int64_t rows_per_chunk = (N + chunks - 1) / chunks;
  rows_per_chunk = (rows_per_chunk + 31) / 32 * 32;
  chunks = (int)std::max<int64_t>(1, (N + rows_per_chunk - 1) / rows_per_chunk);

  float *part_d;
  uint32_t *part_i;
  int rc;
  if ((rc = annb_scratch(h, 16, (size_t)B * chunks * k * sizeof(float), (void **)&part_d))) return rc;
  if ((rc = annb_scratch(h, 17, (size_t)B * chunks * k * sizeof(uint32_t), (void **)&part_i))) return rc;

  const int epl = (k + 31) / 32;
#define ST_CASE(E)                                                                                                    \
  if (h->code_bytes == 1)                                                                                             \
    rc = run_scan_topk<E, uint8_t>(h, d_tables, B, k, chunks, rows_per_chunk, W, in_smem, smem, part_d, part_i);      \
  else                                                                                                                \
    rc = run_scan_topk<E, uint16_t>(h, d_tables, B, k, chunks, rows_per_chunk, W, in_smem, smem, part_d, part_i);
  if (epl <= 1) {
    ST_CASE(1)
  } else if (epl <= 2) {
    ST_CASE(2)
  } else if (epl <= 4) {
    ST_CASE(4)
  } else if (epl <= 8) {
    ST_CASE(8)
  } else {
    ST_CASE(16)
  }
#undef ST_CASE
  if (rc) return rc;
  // merge chunk partials: lists laid out [b][chunk][k] => g_stride = k, b_stride = chunks*k
  const int warps = 4;
  merge_topk_kernel<uint32_t, int64_t, 1><<<(unsigned)((B + warps - 1) / warps), warps * 32, 0, h->stream>>>(
      part_d, part_i, chunks, B, k, (int64_t)k, (int64_t)chunks * k, d_dists, d_ids, LIST_EMPTY_VAL, (int64_t)-1);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

// Shard merge (container.py:130-138 with the (dist, label) order): G ascending lists of k per query, given as two
// strided arrays so that the all-gathered PACKED per-rank buffers ([B*k dists][B*k labels] per rank, one NCCL
// all-gather) are merged in place.  One warp per query: the G*k pairs are staged in shared memory, then every
// pair finds its global rank = its own position + sum over the other lists of a binary search (lists are sorted),
// O(G*k*G*log k / 32) per query instead of the all-pairs count.
__global__ void merge_sorted_kernel(const float *__restrict__ d_in, const uint64_t *__restrict__ l_in, int G, int64_t B, int k,
                                    int64_t d_gstride, int64_t l_gstride, float *__restrict__ d_out,
                                    uint64_t *__restrict__ l_out) {
  extern __shared__ __align__(16) unsigned char msm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * nwarps + warp;
  if (b >= B) return;
  const int total = G * k;
  uint64_t *sl = reinterpret_cast<uint64_t *>(msm) + (size_t)warp * total;
  float *sd = reinterpret_cast<float *>(msm + (size_t)nwarps * total * 8) + (size_t)warp * total;
  for (int x = lane; x < total; x += 32) {
    const int g = x / k, p = x - g * k;
    sd[x] = d_in[g * d_gstride + b * k + p];
    sl[x] = l_in[g * l_gstride + b * k + p];
  }
  __syncwarp();
  int valid = 0;
  for (int x = lane; x < total; x += 32) {
    const uint64_t lx = sl[x];
    if (lx == UINT64_MAX) continue;  // a shard that found fewer than k
    valid++;
    const int gx = x / k;
    const float dx = sd[x];
    int rank = x - gx * k;  // pairs before it in its own (sorted) list
    for (int g = 0; g < G; g++) {
      if (g == gx) continue;
      // number of pairs of list g that precede (dx, lx): order (dist, label), equal pairs by list index
      int lo = 0, hi = k;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float dm = sd[g * k + mid];
        const uint64_t lm = sl[g * k + mid];
        const bool before = lm != UINT64_MAX && (dm < dx || (dm == dx && (lm < lx || (lm == lx && g < gx))));
        if (before) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < k) {
      d_out[b * k + rank] = dx;
      l_out[b * k + rank] = lx;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(FULL_MASK, valid, o);
  for (int p = valid + lane; p < k; p += 32) {
    d_out[b * k + p] = CUDART_INF_F;
    l_out[b * k + p] = UINT64_MAX;
  }
}

int launch_merge_topk(annb_index *h, const uint64_t *labels, const float *dists, int G, int64_t B, int k, int64_t l_gstride,
                      int64_t d_gstride, uint64_t *labels_out, float *dists_out, cudaStream_t stream) {
  if (B == 0) return ANNB_OK;
  const size_t per_warp = (size_t)G * k * 12;
  if (per_warp > 200 * 1024) ANNB_FAIL(ANNB_ELIMIT, "G*k=%d is too large for the shard merge", G * k);
  int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (96 * 1024) / per_warp));
  const size_t smem = warps * per_warp;
  if (smem > 48 * 1024) ANNB_CUDA(cudaFuncSetAttribute(merge_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_sorted_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, smem, stream>>>(dists, labels, G, B, k, d_gstride, l_gstride,
                                                                                       dists_out, labels_out);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}
