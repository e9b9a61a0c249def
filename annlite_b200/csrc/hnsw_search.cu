// hnsw_search.cu -- K3: warp-per-query HNSW walk over PQ codes (sm_100a).
//
// Reference semantics: HierarchicalNSW::searchKnn (include/hnswlib/hnswalg.h:1237-1295) =
// greedy descent on levels maxlevel..1 (:1248-1274) + searchBaseLayerST (:243-329), and the
// filtered twin searchKnnWithFilter (:1297-1361) + searchBaseLayerSTWithFilter (:332-440), with
// the PQ distance PQLookup (include/hnswlib/space_pq.h:16-37): d = sum_m T[m][code_m], m
// sequential from 0.f -- one lane scores one neighbour, so the fp32 sum order is the reference's.
//
// Kernels:
//  * hnsw_walk_fast     -- no deletions, no filter, maxM0 <= 32 (the headline path).  The reference's two
//    heaps collapse into ONE sorted list of ef entries (registers + a shared-memory mirror): without
//    deletions every candidate is also a top-candidate, so the candidate heap is exactly the not-yet-
//    expanded members of the list (a candidate evicted from top_candidates has d >= lowerBound and the
//    loop would break before expanding it, :270).  No visited set is needed either: a re-encountered
//    node is either still in the list (dropped by an id compare restricted to equal keys) or was
//    rejected / evicted with d >= lowerBound, and lowerBound only falls once the list is full, so it is
//    rejected again (:306).  The walk therefore WRITES NOTHING to global memory until its k results.
//    Per hop: score all <= 32 neighbours (one lane each), rank the admitted ones by a branch-free
//    binary search over the mirror, request the likely next node's record (overlaps the merge), merge
//    by rank counting + scatter through the mirror, then read the true next node off the merged list.
//    Differences from the reference are confined to exact fp32 distance ties.
//  * hnsw_walk_chunked  -- same semantics for graphs with maxM0 > 32 (neighbour list in 32-wide chunks,
//    WarpList::merge per chunk, no prefetch).
//  * hnsw_walk_general  -- deletions and/or filter: follows the reference literally with a separate
//    candidate bag (shared memory, spilling to per-warp global scratch) and an exact visited bitmap
//    (per-warp global scratch, L2 resident), because nodes that fail the filter are traversed but never
//    admitted, so the single-list argument above does not hold.
//
// Memory layout: node record = [maxM0 links][maxM0 neighbour codes], so one
// hop is one contiguous, coalesced read (384 B for M=8) instead of an adjacency read followed by
// up to 32 dependent 8-byte gathers.  The per-query table (M*Ks fp32) is staged into shared memory
// by one TMA bulk copy (cp.async.bulk + mbarrier) per query.
#include <math_constants.h>

#include <algorithm>

#include "annb_internal.h"
#include "tma_utils.cuh"
#include "warp_list.cuh"

namespace {

constexpr uint32_t EMPTY_LINK = 0xffffffffu;
constexpr uint32_t EXPANDED_BIT = 0x80000000u;
constexpr uint32_t ID_MASK = 0x7fffffffu;

// ---- code row in registers -------------------------------------------------------------------
template <int CR>
struct CodeWords {
  uint32_t w[CR / 4];
  __device__ __forceinline__ void load(const uint8_t *p) {
    if (CR == 4) {
      w[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
    } else if (CR == 8) {
      uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
      w[0] = t.x;
      w[1] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < CR / 16; i++) {
        uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        w[4 * i + 0] = t.x;
        w[4 * i + 1] = t.y;
        w[4 * i + 2] = t.z;
        w[4 * i + 3] = t.w;
      }
    }
  }
};

// PQLookup (space_pq.h:30-35): strictly sequential fp32 sum over subquantisers
template <int CR, int CB>
__device__ __forceinline__ float pq_lookup(const float *T, const CodeWords<CR> &c, int Ks) {
  float r = 0.f;
  constexpr int M = CR / CB;
#pragma unroll
  for (int m = 0; m < M; m++) {
    uint32_t code;
    if (CB == 1) code = (c.w[m >> 2] >> (8 * (m & 3))) & 0xffu;
    else code = (c.w[m >> 1] >> (16 * (m & 1))) & 0xffffu;
    r = __fadd_rn(r, T[m * Ks + code]);
  }
  return r;
}
// generic: code bytes straight from memory
template <int CB>
__device__ __forceinline__ float pq_lookup_mem(const float *T, const uint8_t *p, int M, int Ks) {
  float r = 0.f;
  for (int m = 0; m < M; m++) {
    uint32_t code = CB == 1 ? (uint32_t)p[m] : (uint32_t)reinterpret_cast<const uint16_t *>(p)[m];
    r = __fadd_rn(r, T[m * Ks + code]);
  }
  return r;
}

template <int CR, int CB>
__device__ __forceinline__ float score(const float *T, const uint8_t *code_ptr, int M, int Ks) {
  if (CR > 0) {
    CodeWords<(CR > 0 ? CR : 4)> c;
    c.load(code_ptr);
    return pq_lookup<(CR > 0 ? CR : 4), CB>(T, c, Ks);
  } else {
    return pq_lookup_mem<CB>(T, code_ptr, M, Ks);
  }
}

__device__ __forceinline__ void load_table(float *dst, const float *__restrict__ src, int TS, int lane) {
  if ((TS & 3) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int i = lane; i < (TS >> 2); i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = lane; i < TS; i += 32) dst[i] = __ldg(src + i);
  }
}

// Next query of the persistent loop: ONE lane bumps the work counter (a predicated instruction, not a branch) and
// the value is spread with a warp REDUCTION: REDUX delivers it in a uniform register, so ptxas can prove that the
// loop exit is warp-uniform.  With `__shfl_sync(qi, 0)` it cannot, and then guards every warp-level operation of the
// loop body with a BRA.DIV convergence check.
__device__ __forceinline__ int64_t next_query(unsigned int *counter, int lane) {
  unsigned qi = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.u32 p, %2, 0;\n"
      "@p atom.global.add.u32 %0, [%1], 1;\n"
      "}\n"
      : "+r"(qi)
      : "l"(counter), "r"(lane)
      : "memory");
  return (int64_t)__reduce_max_sync(FULL_MASK, qi);
}

struct Walk {
  uint32_t node;   // node reached on level 0
  float dist;      // its distance
  int hops, nbrs, evals;
};

// Greedy descent maxlevel..1 (hnswalg.h:1245-1274).  After scanning one node's list the
// reference holds the FIRST minimum among neighbours that beat curdist (strict <, sequential).
template <int CR, int CB>
__device__ __forceinline__ Walk descend(const GraphDev &g, const float *T, int lane) {
  Walk w;
  w.hops = 0;
  w.nbrs = 0;
  w.evals = 1;
  w.dist = pq_lookup_mem<CB>(T, g.ep_code, g.M, g.Ks);  // dist to the entry point (:1246)
  uint32_t rec = g.ep_rec;
  for (int level = g.maxlevel; level > 0; level--) {
    const uint8_t *base = g.up + g.up_off[level];
    bool changed = true;
    while (changed) {
      changed = false;
      const uint8_t *r = base + (size_t)rec * g.recu_bytes;
      float best = w.dist;
      uint32_t best_link = EMPTY_LINK;
      w.hops++;
      for (int c0 = 0; c0 < g.maxM; c0 += 32) {
        const int j = c0 + lane;
        uint32_t link = j < g.maxM ? __ldg(reinterpret_cast<const uint32_t *>(r) + j) : EMPTY_LINK;
        const bool valid = link != EMPTY_LINK;
        float d = CUDART_INF_F;
        if (valid) d = score<CR, CB>(T, r + g.code_offu + (size_t)j * g.code_row, g.M, g.Ks);
        const int nv = __popc(__ballot_sync(FULL_MASK, valid));
        w.nbrs += nv;
        w.evals += nv;
        // warp argmin, first index wins ties
        float md = d;
        int mj = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          float od = __shfl_xor_sync(FULL_MASK, md, o);
          int oj = __shfl_xor_sync(FULL_MASK, mj, o);
          if (od < md || (od == md && oj < mj)) {
            md = od;
            mj = oj;
          }
        }
        if (md < best) {
          best = md;
          best_link = __shfl_sync(FULL_MASK, link, mj);
        }
        if (nv < 32) break;
      }
      if (best_link != EMPTY_LINK) {
        w.dist = best;
        rec = best_link;
        changed = true;
      }
    }
    // step down: record index on the level below (node id when level == 1)
    rec = __ldg(reinterpret_cast<const uint32_t *>(base + (size_t)rec * g.recu_bytes + g.tail_offu) + 1);
  }
  w.node = rec;
  return w;
}

template <int EPL>
__device__ __forceinline__ void write_results(const GraphDev &g, const SearchParams &p, int64_t q, const WarpList<EPL> &L,
                                              int lane, int hops, int nbrs, int evals) {
  const int k = p.k;
  bool tie = false;
  int found = 0;
#pragma unroll
  for (int e = 0; e < EPL; e++) {
    const int pos = e * 32 + lane;
    const bool have = L.v[e] != LIST_EMPTY_VAL && pos < k;
    found += __popc(__ballot_sync(FULL_MASK, have));
    // tie detection between adjacent list positions
    float nk = __shfl_down_sync(FULL_MASK, L.k[e], 1);
    const float first_next = (e + 1 < EPL) ? __shfl_sync(FULL_MASK, L.k[(e + 1 < EPL) ? e + 1 : e], 0) : CUDART_INF_F;
    if (lane == 31) nk = first_next;
    tie |= have && (pos + 1 < k) && (nk == L.k[e]);
    if (pos < k) {
      p.out_dists[q * k + pos] = have ? L.k[e] : CUDART_INF_F;
      p.out_labels[q * k + pos] = have ? __ldg(g.labels + (L.v[e] & ID_MASK)) : (uint64_t)UINT64_MAX;
    }
  }
  tie = __any_sync(FULL_MASK, tie);
  if (tie) {
    // rows must be ascending by (dist, label) (hnsw_bindings.cpp:346-351); equal distances are
    // rare, fix them up serially
    __syncwarp();
    if (lane == 0) {
      for (int i = 1; i < found; i++) {
        float d = p.out_dists[q * k + i];
        uint64_t l = p.out_labels[q * k + i];
        int j = i - 1;
        while (j >= 0 && p.out_dists[q * k + j] == d && p.out_labels[q * k + j] > l) {
          p.out_dists[q * k + j + 1] = p.out_dists[q * k + j];
          p.out_labels[q * k + j + 1] = p.out_labels[q * k + j];
          j--;
        }
        p.out_dists[q * k + j + 1] = d;
        p.out_labels[q * k + j + 1] = l;
      }
    }
  }
  if (lane == 0) {
    p.out_found[q] = found;
    if (p.out_stats) {
      p.out_stats[q * 3 + 0] = hops;
      p.out_stats[q * 3 + 1] = nbrs;
      p.out_stats[q * 3 + 2] = evals;
    }
  }
}

// =================================================================================================
// fast path: no deletions, no filter
// =================================================================================================
template <int EPL, int CR, int CB, bool SMEM_TABLE>
__global__ void hnsw_walk_chunked(const GraphDev g, const SearchParams p, const int table_stride_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int TS = g.M * g.Ks;
  // [nwarps x table (SMEM_TABLE only)] [nwarps x 32*EPL uint2 merge scratch]
  float *Ts = reinterpret_cast<float *>(smem_raw + (size_t)warp * table_stride_bytes);
  uint2 *scratch = reinterpret_cast<uint2 *>(smem_raw + (size_t)nwarps * table_stride_bytes) + (size_t)warp * 32 * EPL;
  const int ef = p.ef;

  for (;;) {
    const int64_t q = next_query(p.work_counter, lane);
    if (q >= p.B) break;
    const float *T;
    if (SMEM_TABLE) {
      __syncwarp();
      load_table(Ts, p.tables + q * TS, TS, lane);
      __syncwarp();
      T = Ts;
    } else {
      T = p.tables + q * TS;
    }

    Walk w = descend<CR, CB>(g, T, lane);
    int hops = w.hops, nbrs = w.nbrs, evals = w.evals + 1;  // searchBaseLayerST re-scores the entry (:255)

    WarpList<EPL> L;
    L.clear();
    int size = 0;
    float worst = CUDART_INF_F;  // lowerBound; +inf while the list is not full (:306)
    L.merge(w.dist, w.node, lane == 0, ef, scratch, size, worst, ID_MASK);

    for (;;) {
      // nearest not-yet-expanded list entry == candidate_set.top() (:268)
      int pos = -1;
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const unsigned m = __ballot_sync(FULL_MASK, !(L.v[e] & EXPANDED_BIT));  // empty slots have the bit set
        if (m && pos < 0) pos = e * 32 + __ffs(m) - 1;
      }
      if (pos < 0) break;
      const uint32_t node = L.val_at(pos);
#pragma unroll
      for (int e = 0; e < EPL; e++)
        if (e * 32 + lane == pos) L.v[e] |= EXPANDED_BIT;

      const uint8_t *rec = g.rec0 + (size_t)node * g.rec0_bytes;
      hops++;
      for (int c0 = 0; c0 < g.maxM0; c0 += 32) {
        const int j = c0 + lane;
        const uint32_t link = j < g.maxM0 ? __ldg(reinterpret_cast<const uint32_t *>(rec) + j) : EMPTY_LINK;
        const bool valid = link != EMPTY_LINK;
        float d = CUDART_INF_F;
        if (valid) d = score<CR, CB>(T, rec + g.code_off0 + (size_t)j * g.code_row, g.M, g.Ks);
        const int nv = __popc(__ballot_sync(FULL_MASK, valid));
        nbrs += nv;
        evals += nv;
        // admission test of :306 for the whole neighbour list at once; candidates that a sequential
        // scan would reject after lowerBound fell land beyond position ef-1 in the merge and drop out
        L.merge(d, link, valid && d < worst, ef, scratch, size, worst, ID_MASK);
        if (nv < 32) break;
      }
    }
    write_results<EPL>(g, p, q, L, lane, hops, nbrs, evals);
  }
}

// ---- v3: one neighbour chunk per node (maxM0 <= 32, the reference default M=16) ------------------
// Software-pipelined hop: as soon as the neighbours are scored the most likely next node is known
// (the nearest unexpanded list entry, or a new candidate that beats it), so its 384 B record is
// requested BEFORE the merge and the DRAM round trip overlaps the merge work.  The next node is then
// determined exactly from the merged list; a wrong guess only costs a reload.
template <int EPL, int CR, int CB, bool SMEM_TABLE>
__global__ void hnsw_walk_fast(const GraphDev g, const SearchParams p, const int table_stride_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int TS = g.M * g.Ks;
  constexpr int CAP = 32 * EPL;
  // [nwarps x table (SMEM_TABLE only)] [nwarps x (CAP list mirror + 32 candidate slots) uint2]
  float *Ts = reinterpret_cast<float *>(smem_raw + (size_t)warp * table_stride_bytes);
  // [nwarps x table] [nwarps x (CAP list mirror + 32 candidate slots + 1 mbarrier) x 8 B]
  uint2 *sl = reinterpret_cast<uint2 *>(smem_raw + (size_t)nwarps * table_stride_bytes) + (size_t)warp * (CAP + 33);
  uint2 *cbuf = sl + CAP;
  uint64_t *tbar = reinterpret_cast<uint64_t *>(cbuf + 32);
  const bool tma_table = SMEM_TABLE && ((TS * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.tables) & 15) == 0);
  uint32_t tphase = 0;
  if (tma_table) {
    if (lane == 0) mbar_init(tbar, 1);
    __syncwarp();
  }
  const int ef = p.ef;
  const unsigned lt_mask = (1u << lane) - 1u;
  constexpr int CW = CR > 0 ? CR / 4 : 1;

  for (;;) {
    const int64_t q = next_query(p.work_counter, lane);
    if (q >= p.B) break;
    const float *T;
    if (SMEM_TABLE) {
      __syncwarp();  // every lane is done reading the previous query's table
      if (tma_table) {
        if (lane == 0) {
          mbar_expect_tx(tbar, (uint32_t)TS * 4u);
          bulk_g2s(Ts, p.tables + q * TS, (uint32_t)TS * 4u, tbar);
        }
        mbar_wait(tbar, tphase);
        tphase ^= 1u;
      } else {
        load_table(Ts, p.tables + q * TS, TS, lane);
        __syncwarp();
      }
      T = Ts;
    } else {
      T = p.tables + q * TS;
    }

    Walk w = descend<CR, CB>(g, T, lane);
    int hops = w.hops, nbrs = w.nbrs, evals = w.evals + 1;  // searchBaseLayerST re-scores the entry (:255)

    // list: registers (striped) + shared mirror; the entry node starts expanded (it is hop 0)
    WarpList<EPL> L;
    L.clear();
    // the shared mirror is padded with +inf keys so the rank search below needs no bounds
#pragma unroll
    for (int e = 0; e < EPL; e++) sl[e * 32 + lane] = make_uint2(0x7f800000u, LIST_EMPTY_VAL);
    __syncwarp();
    if (lane == 0) {
      L.k[0] = w.dist;
      L.v[0] = w.node | EXPANDED_BIT;
      sl[0] = make_uint2(__float_as_uint(w.dist), w.node | EXPANDED_BIT);
    }
    int size = 1;
    float worst = (ef == 1) ? w.dist : CUDART_INF_F;  // lowerBound (:306); +inf while the list is not full
    __syncwarp();

    // current record in registers
    uint32_t link;
    uint32_t cw[CW];
    const uint8_t *code_ptr = nullptr;
    const bool lane_has_slot = lane < g.maxM0;
    const uint8_t *lane_link_base = g.rec0 + 4 * lane;
    const uint8_t *lane_code_base = g.rec0 + g.code_off0 + (size_t)lane * g.code_row;
    const uint32_t rec_bytes = (uint32_t)g.rec0_bytes;
    auto load_record = [&](uint32_t node, uint32_t &lk, uint32_t *words, const uint8_t *&cptr) {
      const size_t off = (size_t)node * rec_bytes;
      lk = lane_has_slot ? __ldg(reinterpret_cast<const uint32_t *>(lane_link_base + off)) : EMPTY_LINK;
      cptr = lane_code_base + off;
      if (CR > 0 && lane_has_slot) {
        CodeWords<(CR > 0 ? CR : 4)> c;
        c.load(cptr);
#pragma unroll
        for (int i = 0; i < CW; i++) words[i] = c.w[i];
      }
    };
    load_record(w.node, link, cw, code_ptr);

    for (;;) {
      hops++;
      // ---- nearest unexpanded entry of the current list; request its record right away: unless this
      // hop finds something closer it is the next node, and its DRAM round trip then overlaps the
      // whole hop (scoring + merge) ----
      int pos2 = -1;
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const unsigned m = __ballot_sync(FULL_MASK, !(L.v[e] & EXPANDED_BIT));
        if (m && pos2 < 0) pos2 = e * 32 + __ffs(m) - 1;
      }
      uint2 e2 = make_uint2(0x7f800000u, LIST_EMPTY_VAL);
      if (pos2 >= 0) e2 = sl[pos2];
      uint32_t pred = LIST_EMPTY_VAL;
      uint32_t link_n = EMPTY_LINK;
      uint32_t cw_n[CW];
      const uint8_t *code_ptr_n = nullptr;
      if (pos2 >= 0) {
        pred = e2.y & ID_MASK;
        load_record(pred, link_n, cw_n, code_ptr_n);
      }
      // ---- score the neighbour list (one lane = one neighbour, m sequential) ----
      const bool valid = link != EMPTY_LINK;
      float d = CUDART_INF_F;
      if (valid) {
        if (CR > 0) {
          CodeWords<(CR > 0 ? CR : 4)> c;
#pragma unroll
          for (int i = 0; i < CW; i++) c.w[i] = cw[i];
          d = pq_lookup<(CR > 0 ? CR : 4), CB>(T, c, g.Ks);
        } else {
          d = pq_lookup_mem<CB>(T, code_ptr, g.M, g.Ks);
        }
      }
      const int nv = __popc(__ballot_sync(FULL_MASK, valid));
      nbrs += nv;
      evals += nv;

      // ---- admission (:306) for the whole list at once: rank by binary search, drop re-encounters ----
      const bool take = valid && d < worst;
      const unsigned offered = __ballot_sync(FULL_MASK, take);
      unsigned live = 0;
      int base = 0;
      if (offered) {
        bool dup = false;
        if (take) {
          // branch-free lower bound over the CAP-entry (inf padded) mirror; d < worst <= key[ef-1]
          // guarantees the count fits in [0, CAP-1]
          int lo = 0;
#pragma unroll
          for (int step = CAP / 2; step >= 1; step >>= 1) lo += (__uint_as_float(sl[lo + step - 1].x) <= d) ? step : 0;
          base = lo;  // number of list keys <= d: the candidate lands after its equals
          for (int t = base - 1; t >= 0; t--) {  // an id can only match where the key matches
            const uint2 x = sl[t];
            if (__uint_as_float(x.x) != d) break;
            if ((x.y & ID_MASK) == link) {
              dup = true;
              break;
            }
          }
        }
        live = __ballot_sync(FULL_MASK, take && !dup);
      }

      // ---- a new candidate beats the entry whose record is already on its way: re-aim the prefetch ----
      const unsigned fm = __ballot_sync(FULL_MASK, ((live >> lane) & 1u) && d < __uint_as_float(e2.x));
      if (fm) {
        pred = __shfl_sync(FULL_MASK, link, __ffs(fm) - 1);
        load_record(pred, link_n, cw_n, code_ptr_n);
      }

      // ---- merge the live candidates into the list ----
      int pos = pos2;
      if (live) {
        const int r = __popc(live & lt_mask);
        const bool mine = (live >> lane) & 1u;
        if (mine) cbuf[r] = make_uint2(__float_as_uint(d), link);
        __syncwarp();
        const int np_live = __popc(live);
        int shift[EPL];
#pragma unroll
        for (int e = 0; e < EPL; e++) shift[e] = 0;
        int crank = 0;
        for (int i = 0; i < np_live; i++) {
          const float dc = __uint_as_float(cbuf[i].x);
#pragma unroll
          for (int e = 0; e < EPL; e++) shift[e] += (dc < L.k[e]) ? 1 : 0;
          crank += ((dc < d) || (dc == d && i < r)) ? 1 : 0;
        }
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int np = e * 32 + lane + shift[e];
          if (L.v[e] != LIST_EMPTY_VAL && np < ef) sl[np] = make_uint2(__float_as_uint(L.k[e]), L.v[e]);
        }
        if (mine) {
          const int np = base + crank;
          if (np < ef) sl[np] = make_uint2(__float_as_uint(d), link);
        }
        size = min(size + np_live, ef);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int ps = e * 32 + lane;
          if (ps < size) {
            const uint2 t = sl[ps];
            L.k[e] = __uint_as_float(t.x);
            L.v[e] = t.y;
          } else {
            L.k[e] = CUDART_INF_F;
            L.v[e] = LIST_EMPTY_VAL;
          }
        }
        if (size == ef) worst = __uint_as_float(sl[ef - 1].x);
        // no live candidate beats the old nearest-unexpanded entry => it did not move and is still first
        if (fm || pos2 < 0) {
          pos = -1;
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const unsigned m = __ballot_sync(FULL_MASK, !(L.v[e] & EXPANDED_BIT));
            if (m && pos < 0) pos = e * 32 + __ffs(m) - 1;
          }
        }
      }
      if (pos < 0) break;  // candidate_set exhausted (:266)

      // ---- expand `pos`: flag it in both copies, take its record ----
      const uint32_t node = sl[pos].y & ID_MASK;
#pragma unroll
      for (int e = 0; e < EPL; e++)
        if (e * 32 + lane == pos) {
          L.v[e] |= EXPANDED_BIT;
          sl[pos].y = L.v[e];
        }
      __syncwarp();
      if (node == pred) {
        link = link_n;
        code_ptr = code_ptr_n;
#pragma unroll
        for (int i = 0; i < CW; i++) cw[i] = cw_n[i];
      } else {
        load_record(node, link, cw, code_ptr);
      }
    }
    write_results<EPL>(g, p, q, L, lane, hops, nbrs, evals);
  }
}

// =================================================================================================
// flagged path: deletions and/or filter WITHOUT a visited set or a candidate bag
// =================================================================================================
// Same single-list idea as hnsw_walk_fast, extended to nodes that are traversed but not admitted:
// every evaluated node that the reference would push to candidate_set (:306 / :413) is kept in ONE
// sorted list (capacity CAP = 32*EPL, chosen by the host from the filter's selectivity) with a PASS
// flag (admitted to top_candidates: passes the filter / is not deleted).  lowerBound is read off the
// list: the key of the ef-th passing entry once ef of them exist (everything behind it is dropped --
// those candidates have d >= lowerBound and are never expanded, :270 / :371), else the key of the last
// passing entry, else FLT_MAX -- exactly the values top_candidates.top() takes in the reference.
// While fewer than ef entries pass nothing is dropped, which is what makes the walk exact without a
// visited set: a re-encountered node is either still listed (id compare on equal keys) or was dropped
// with d >= a lowerBound that has only fallen since.  If the list would overflow its capacity the
// query is flagged (found = -1) and the host re-runs the batch on the bitmap kernel below.
constexpr uint32_t PASS_BIT = 0x40000000u;
constexpr uint32_t ID_MASK30 = 0x3fffffffu;

template <int EPL, int CR, int CB>
__global__ void hnsw_walk_flagged(const GraphDev g, const SearchParams p, const int has_del, const int table_stride_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int TS = g.M * g.Ks;
  constexpr int CAP = 32 * EPL;
  float *Ts = reinterpret_cast<float *>(smem_raw + (size_t)warp * table_stride_bytes);
  uint2 *sl = reinterpret_cast<uint2 *>(smem_raw + (size_t)nwarps * table_stride_bytes) + (size_t)warp * (CAP + 33);
  uint2 *cbuf = sl + CAP;
  uint64_t *tbar = reinterpret_cast<uint64_t *>(cbuf + 32);
  const bool tma_table = ((TS * 4) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.tables) & 15) == 0);
  uint32_t tphase = 0;
  if (tma_table) {
    if (lane == 0) mbar_init(tbar, 1);
    __syncwarp();
  }
  const int ef = p.ef;
  const int k = p.k;
  const unsigned lt_mask = (1u << lane) - 1u;
  const unsigned le_mask = lt_mask | (1u << lane);
  constexpr int CW = CR / 4;
  const uint32_t *filter = p.filter;
  const bool use_filter = filter != nullptr;
  const uint32_t *delbits = g.deleted;
  auto passes = [&](uint32_t id) -> bool {
    if (use_filter) return (__ldg(filter + (id >> 5)) >> (id & 31)) & 1u;   // :353-354, :423-426
    return !((__ldg(delbits + (id >> 5)) >> (id & 31)) & 1u);              // :254, :314
  };
  const float FLT_MAX_ = 3.402823466e+38f;

  for (;;) {
    const int64_t q = next_query(p.work_counter, lane);
    if (q >= p.B) break;
    __syncwarp();
    if (tma_table) {
      if (lane == 0) {
        mbar_expect_tx(tbar, (uint32_t)TS * 4u);
        bulk_g2s(Ts, p.tables + q * TS, (uint32_t)TS * 4u, tbar);
      }
      mbar_wait(tbar, tphase);
      tphase ^= 1u;
    } else {
      load_table(Ts, p.tables + q * TS, TS, lane);
      __syncwarp();
    }
    const float *T = Ts;

    Walk w = descend<CR, CB>(g, T, lane);
    int hops = w.hops, nbrs = w.nbrs, evals = w.evals + 1;

    WarpList<EPL> L;
    L.clear();
#pragma unroll
    for (int e = 0; e < EPL; e++) sl[e * 32 + lane] = make_uint2(0x7f800000u, LIST_EMPTY_VAL);
    __syncwarp();
    const bool ep_pass = passes(w.node);
    const uint32_t ep_val = w.node | EXPANDED_BIT | (ep_pass ? PASS_BIT : 0u);
    if (lane == 0) {
      L.k[0] = w.dist;
      L.v[0] = ep_val;
      sl[0] = make_uint2(__float_as_uint(w.dist), ep_val);
    }
    int size = 1, npass = ep_pass ? 1 : 0;
    float lb = ep_pass ? w.dist : FLT_MAX_;   // lowerBound of the reference (:256 / :260)
    bool aborted = false;
    __syncwarp();

    uint32_t link;
    uint32_t cw[CW];
    const bool lane_has_slot = lane < g.maxM0;
    const uint8_t *lane_link_base = g.rec0 + 4 * lane;
    const uint8_t *lane_code_base = g.rec0 + g.code_off0 + (size_t)lane * g.code_row;
    const uint32_t rec_bytes = (uint32_t)g.rec0_bytes;
    auto load_record = [&](uint32_t node, uint32_t &lk, uint32_t *words) {
      const size_t off = (size_t)node * rec_bytes;
      lk = lane_has_slot ? __ldg(reinterpret_cast<const uint32_t *>(lane_link_base + off)) : EMPTY_LINK;
      if (lane_has_slot) {
        CodeWords<CR> c;
        c.load(lane_code_base + off);
#pragma unroll
        for (int i = 0; i < CW; i++) words[i] = c.w[i];
      }
    };
    load_record(w.node, link, cw);

    for (;;) {
      hops++;
      int pos2 = -1;
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        const unsigned m = __ballot_sync(FULL_MASK, !(L.v[e] & EXPANDED_BIT));
        if (m && pos2 < 0) pos2 = e * 32 + __ffs(m) - 1;
      }
      uint2 e2 = make_uint2(0x7f800000u, LIST_EMPTY_VAL);
      if (pos2 >= 0) e2 = sl[pos2];
      uint32_t pred = LIST_EMPTY_VAL;
      uint32_t link_n = EMPTY_LINK;
      uint32_t cw_n[CW];
      if (pos2 >= 0) {  // the likely next node: its record travels while this hop is scored and merged
        pred = e2.y & ID_MASK30;
        load_record(pred, link_n, cw_n);
      }
      const bool valid = link != EMPTY_LINK;
      float d = CUDART_INF_F;
      bool pf = false;
      if (valid) {
        pf = passes(link);
        CodeWords<CR> c;
#pragma unroll
        for (int i = 0; i < CW; i++) c.w[i] = cw[i];
        d = pq_lookup<CR, CB>(T, c, g.Ks);
      }
      const int nv = __popc(__ballot_sync(FULL_MASK, valid));
      nbrs += nv;
      evals += nv;

      // admission to candidate_set: top_candidates.size() < ef || lowerBound > dist (:306 / :413)
      const bool take = valid && (npass < ef || d < lb);
      const unsigned offered = __ballot_sync(FULL_MASK, take);
      unsigned live = 0;
      int base = 0;
      if (offered) {
        bool dup = false;
        if (take) {
          int lo = 0;
#pragma unroll
          for (int step = CAP / 2; step >= 1; step >>= 1) lo += (__uint_as_float(sl[lo + step - 1].x) <= d) ? step : 0;
          base = lo;
          for (int t = base - 1; t >= 0; t--) {
            const uint2 x = sl[t];
            if (__uint_as_float(x.x) != d) break;
            if ((x.y & ID_MASK30) == link) {
              dup = true;
              break;
            }
          }
        }
        live = __ballot_sync(FULL_MASK, take && !dup);
      }
      const int np_live = __popc(live);
      if (size + np_live > CAP) {  // capacity chosen by the host was too small for this query
        aborted = true;
        break;
      }

      const unsigned fm = __ballot_sync(FULL_MASK, ((live >> lane) & 1u) && d < __uint_as_float(e2.x));
      if (fm) {
        pred = __shfl_sync(FULL_MASK, link, __ffs(fm) - 1);
        load_record(pred, link_n, cw_n);
      }

      int pos = pos2;
      if (live) {
        const int r = __popc(live & lt_mask);
        const bool mine = (live >> lane) & 1u;
        const uint32_t myval = link | (pf ? PASS_BIT : 0u);
        if (mine) cbuf[r] = make_uint2(__float_as_uint(d), myval);
        __syncwarp();
        int shift[EPL];
#pragma unroll
        for (int e = 0; e < EPL; e++) shift[e] = 0;
        int crank = 0;
        for (int i = 0; i < np_live; i++) {
          const float dc = __uint_as_float(cbuf[i].x);
#pragma unroll
          for (int e = 0; e < EPL; e++) shift[e] += (dc < L.k[e]) ? 1 : 0;
          crank += ((dc < d) || (dc == d && i < r)) ? 1 : 0;
        }
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int np = e * 32 + lane + shift[e];
          if (L.v[e] != LIST_EMPTY_VAL) sl[np] = make_uint2(__float_as_uint(L.k[e]), L.v[e]);
        }
        if (mine) sl[base + crank] = make_uint2(__float_as_uint(d), myval);
        const int merged = size + np_live;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int ps = e * 32 + lane;
          if (ps < merged) {
            const uint2 t = sl[ps];
            L.k[e] = __uint_as_float(t.x);
            L.v[e] = t.y;
          } else {
            L.k[e] = CUDART_INF_F;
            L.v[e] = LIST_EMPTY_VAL;
          }
        }
        // ---- lowerBound = top_candidates.top() (:320-321) read off the list; drop what lies beyond it ----
        unsigned pm[EPL];
        int total = 0;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          pm[e] = __ballot_sync(FULL_MASK, L.v[e] != LIST_EMPTY_VAL && (L.v[e] & PASS_BIT));
          total += __popc(pm[e]);
        }
        size = merged;
        if (total >= ef) {
          int cum = 0, P = -1;
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int c = __popc(pm[e]);
            if (P < 0 && cum + c >= ef) {
              const int need = ef - cum;
              const unsigned hit = __ballot_sync(FULL_MASK, ((pm[e] >> lane) & 1u) && __popc(pm[e] & le_mask) == need);
              P = e * 32 + __ffs(hit) - 1;
            }
            cum += c;
          }
          const int newsize = P + 1;
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const int ps = e * 32 + lane;
            if (ps >= newsize && ps < merged) {
              L.k[e] = CUDART_INF_F;
              L.v[e] = LIST_EMPTY_VAL;
              sl[ps] = make_uint2(0x7f800000u, LIST_EMPTY_VAL);
            }
          }
          size = newsize;
          npass = ef;
          lb = __uint_as_float(sl[P].x);
        } else {
          npass = total;
          if (total > 0) {
            int plast = -1;
#pragma unroll
            for (int e = EPL - 1; e >= 0; e--)
              if (plast < 0 && pm[e]) plast = e * 32 + 31 - __clz(pm[e]);
            lb = __uint_as_float(sl[plast].x);
          }
        }
        __syncwarp();
        pos = -1;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const unsigned m = __ballot_sync(FULL_MASK, !(L.v[e] & EXPANDED_BIT));
          if (m && pos < 0) pos = e * 32 + __ffs(m) - 1;
        }
      }
      if (pos < 0) break;
      const uint2 nx = sl[pos];
      const float nkey = __uint_as_float(nx.x);
      if (use_filter) {
        if (nkey > lb) break;                                                      // :371
      } else if (nkey > lb && (npass == ef || !has_del)) {
        break;                                                                     // :270
      }
      const uint32_t node = nx.y & ID_MASK30;
#pragma unroll
      for (int e = 0; e < EPL; e++)
        if (e * 32 + lane == pos) {
          L.v[e] |= EXPANDED_BIT;
          sl[pos].y = L.v[e];
        }
      __syncwarp();
      if (node == pred) {
        link = link_n;
#pragma unroll
        for (int i = 0; i < CW; i++) cw[i] = cw_n[i];
      } else {
        load_record(node, link, cw);
      }
    }

    // ---- results: the first k passing entries, ascending (dist, label) ----
    if (aborted) {
      if (lane == 0) p.out_found[q] = -1;
      continue;
    }
    int cum = 0;
    bool tie = false;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const bool pass = L.v[e] != LIST_EMPTY_VAL && (L.v[e] & PASS_BIT);
      const unsigned m = __ballot_sync(FULL_MASK, pass);
      const int r = cum + __popc(m & lt_mask);
      if (pass && r < k) {
        p.out_dists[q * k + r] = L.k[e];
        p.out_labels[q * k + r] = __ldg(g.labels + (L.v[e] & ID_MASK30));
      }
      float nk = __shfl_down_sync(FULL_MASK, L.k[e], 1);
      const float first_next = (e + 1 < EPL) ? __shfl_sync(FULL_MASK, L.k[(e + 1 < EPL) ? e + 1 : e], 0) : CUDART_INF_F;
      if (lane == 31) nk = first_next;
      tie |= (L.v[e] != LIST_EMPTY_VAL) && (nk == L.k[e]);
      cum += __popc(m);
    }
    const int found = min(cum, k);
    for (int r = found + lane; r < k; r += 32) {
      p.out_dists[q * k + r] = CUDART_INF_F;
      p.out_labels[q * k + r] = (uint64_t)UINT64_MAX;
    }
    if (__any_sync(FULL_MASK, tie)) {
      __syncwarp();
      if (lane == 0) {
        for (int i = 1; i < found; i++) {
          float dd = p.out_dists[q * k + i];
          uint64_t ll = p.out_labels[q * k + i];
          int j = i - 1;
          while (j >= 0 && p.out_dists[q * k + j] == dd && p.out_labels[q * k + j] > ll) {
            p.out_dists[q * k + j + 1] = p.out_dists[q * k + j];
            p.out_labels[q * k + j + 1] = p.out_labels[q * k + j];
            j--;
          }
          p.out_dists[q * k + j + 1] = dd;
          p.out_labels[q * k + j + 1] = ll;
        }
      }
    }
    if (lane == 0) {
      p.out_found[q] = found;
      if (p.out_stats) {
        p.out_stats[q * 3 + 0] = hops;
        p.out_stats[q * 3 + 1] = nbrs;
        p.out_stats[q * 3 + 2] = evals;
      }
    }
  }
}

// =================================================================================================
// bitmap path: deletions and/or filter (literal two-structure walk with an exact visited set)
// =================================================================================================
__device__ __forceinline__ uint64_t pack_cand(float d, uint32_t id) { return ((uint64_t)__float_as_uint(d) << 32) | id; }
__device__ __forceinline__ float cand_d(uint64_t c) { return __uint_as_float((uint32_t)(c >> 32)); }
__device__ __forceinline__ uint32_t cand_id(uint64_t c) { return (uint32_t)c; }

constexpr int SBAG = 256;  // candidate bag entries kept in shared memory per warp (spills to global beyond)

template <int EPL, int CR, int CB, bool SMEM_TABLE>
__global__ void hnsw_walk_general(const GraphDev g, const SearchParams p, const int has_del, const int table_stride_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  const int TS = g.M * g.Ks;
  float *Ts = reinterpret_cast<float *>(smem_raw + (size_t)warp * table_stride_bytes);
  // [nwarps x table] [nwarps x (32*EPL merge scratch + SBAG bag) x 8 B]
  uint2 *scratch = reinterpret_cast<uint2 *>(smem_raw + (size_t)nwarps * table_stride_bytes) + (size_t)warp * (32 * EPL + SBAG);
  uint64_t *sbag = reinterpret_cast<uint64_t *>(scratch + 32 * EPL);
  const unsigned lt_mask = (1u << lane) - 1u;
  const int ef = p.ef;
  const int64_t slot = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  uint32_t *vis = p.visited + slot * p.visited_words;
  uint32_t *touched = p.touched + slot * (int64_t)p.touched_cap;
  uint64_t *gbag = p.cand + slot * (int64_t)p.cand_cap;
  const uint32_t *filter = p.filter;
  const bool use_filter = filter != nullptr;

  for (;;) {
    const int64_t wi = next_query(p.work_counter, lane);
    if (wi >= p.B) break;
    const int64_t q = p.qmap ? (int64_t)__ldg(p.qmap + wi) : wi;  // subset re-run: outputs stay indexed by query row
    const float *T;
    if (SMEM_TABLE) {
      __syncwarp();
      load_table(Ts, p.tables + q * TS, TS, lane);
      __syncwarp();
      T = Ts;
    } else {
      T = p.tables + q * TS;
    }

    Walk w = descend<CR, CB>(g, T, lane);
    int hops = w.hops, nbrs = w.nbrs, evals = w.evals;

    WarpList<EPL> L;  // top_candidates: admitted entries only
    L.clear();
    // candidate_set: an unordered bag, first SBAG entries in shared memory, the rest in global scratch
    int topsize = 0, ns = 0, ng = 0, ntouched = 0;
    bool overflow = false;
    float lower;
    {
      const uint32_t ep = w.node;
      bool ok;
      if (use_filter) ok = (filter[ep >> 5] >> (ep & 31)) & 1u;                    // :353-354
      else ok = !has_del || !((g.deleted[ep >> 5] >> (ep & 31)) & 1u);             // :254
      if (ok) {
        evals++;
        lower = w.dist;
        L.insert(w.dist, ep, ef);
        topsize = 1;
        if (lane == 0) sbag[0] = pack_cand(w.dist, ep);
      } else {
        lower = 3.402823466e+38f;  // std::numeric_limits<float>::max()
        if (lane == 0) sbag[0] = pack_cand(lower, ep);
      }
      ns = 1;
      if (lane == 0) {
        vis[ep >> 5] |= 1u << (ep & 31);
        touched[0] = ep;
      }
      ntouched = 1;
      __syncwarp();
    }

    while (ns + ng > 0) {
      // candidate_set.top(): nearest candidate (shared part, then the spilled part)
      float bd = CUDART_INF_F;
      int bi = 0x7fffffff;
      uint32_t bid = 0;
      for (int i = lane; i < ns; i += 32) {
        const uint64_t c = sbag[i];
        const float d = cand_d(c);
        if (d < bd || (d == bd && i < bi)) {
          bd = d;
          bi = i;
          bid = cand_id(c);
        }
      }
      for (int i = lane; i < ng; i += 32) {
        const uint64_t c = gbag[i];
        const float d = cand_d(c);
        if (d < bd || (d == bd && i + SBAG < bi)) {
          bd = d;
          bi = i + SBAG;
          bid = cand_id(c);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(FULL_MASK, bd, o);
        int oi = __shfl_xor_sync(FULL_MASK, bi, o);
        uint32_t oid = __shfl_xor_sync(FULL_MASK, bid, o);
        if (od < bd || (od == bd && oi < bi)) {
          bd = od;
          bi = oi;
          bid = oid;
        }
      }
      if (use_filter) {
        if (bd > lower) break;                                                     // :371
      } else if (bd > lower && (topsize == ef || !has_del)) {
        break;                                                                     // :270
      }
      // request the record now; the bag maintenance below overlaps the DRAM round trip
      const uint8_t *rec = g.rec0 + (size_t)bid * g.rec0_bytes;
      uint32_t link0 = (lane < g.maxM0) ? __ldg(reinterpret_cast<const uint32_t *>(rec) + lane) : EMPTY_LINK;
      // pop: move the last entry of that part into the hole
      if (lane == 0) {
        if (bi < SBAG) sbag[bi] = sbag[ns - 1];
        else gbag[bi - SBAG] = gbag[ng - 1];
      }
      if (bi < SBAG) ns--;
      else ng--;
      __syncwarp();
      // candidates beyond lowerBound are dead once top is full (lowerBound only falls): drop them now and
      // then so the nearest-candidate scan stays short and the shared part does not spill
      if (topsize == ef && ((hops & 3) == 3 || ns > SBAG - 40)) {
        int wpos = 0;
        for (int b0 = 0; b0 < ns; b0 += 32) {
          const int i = b0 + lane;
          uint64_t c = 0;
          bool keep = false;
          if (i < ns) {
            c = sbag[i];
            keep = !(cand_d(c) > lower);
          }
          const unsigned km = __ballot_sync(FULL_MASK, keep);
          __syncwarp();
          if (keep) sbag[wpos + __popc(km & lt_mask)] = c;
          wpos += __popc(km);
          __syncwarp();
        }
        ns = wpos;
        if (ng > 0) {
          wpos = 0;
          for (int b0 = 0; b0 < ng; b0 += 32) {
            const int i = b0 + lane;
            uint64_t c = 0;
            bool keep = false;
            if (i < ng) {
              c = gbag[i];
              keep = !(cand_d(c) > lower);
            }
            const unsigned km = __ballot_sync(FULL_MASK, keep);
            __syncwarp();
            if (keep) gbag[wpos + __popc(km & lt_mask)] = c;
            wpos += __popc(km);
            __syncwarp();
          }
          ng = wpos;
        }
      }

      hops++;
      for (int c0 = 0; c0 < g.maxM0; c0 += 32) {
        const int j = c0 + lane;
        const uint32_t link = (c0 == 0) ? link0 : (j < g.maxM0 ? __ldg(reinterpret_cast<const uint32_t *>(rec) + j) : EMPTY_LINK);
        const bool valid = link != EMPTY_LINK;
        const int nv = __popc(__ballot_sync(FULL_MASK, valid));
        nbrs += nv;
        bool fresh = false;
        if (valid) {
          const uint32_t bit = 1u << (link & 31);
          const uint32_t old = atomicOr(vis + (link >> 5), bit);  // exact visited set (:299-301)
          fresh = !(old & bit);
        }
        const unsigned fmask = __ballot_sync(FULL_MASK, fresh);
        const int nf = __popc(fmask);
        evals += nf;
        if (fresh) {
          const int o = ntouched + __popc(fmask & lt_mask);
          if (o < p.touched_cap) touched[o] = link;
        }
        ntouched += nf;
        if (ntouched > p.touched_cap) overflow = true;
        float d = CUDART_INF_F;
        bool admit = false;
        if (fresh) {
          d = score<CR, CB>(T, rec + g.code_off0 + (size_t)j * g.code_row, g.M, g.Ks);
          if (use_filter) admit = (filter[link >> 5] >> (link & 31)) & 1u;        // :423-426
          else admit = !has_del || !((g.deleted[link >> 5] >> (link & 31)) & 1u);  // :314
        }
        // Admission (:306 / :413) for the whole neighbour list at once.  lowerBound only falls once top
        // is full, so testing against its value at the start of the list admits a superset of what the
        // sequential scan admits; the extras have d >= the final lowerBound, are never expanded (the
        // loop breaks first, :270 / :371) and fall off the top list in the merge -- same results.
        const bool cand_ok = fresh && (topsize < ef || lower > d);
        const unsigned cmask = __ballot_sync(FULL_MASK, cand_ok);
        if (cmask) {
          const int cnt = __popc(cmask);
          const int r = __popc(cmask & lt_mask);
          if (ns + cnt <= SBAG) {
            if (cand_ok) sbag[ns + r] = pack_cand(d, link);
            ns += cnt;
          } else {
            if (cand_ok && ng + r < p.cand_cap) gbag[ng + r] = pack_cand(d, link);
            if (ng + cnt > p.cand_cap) overflow = true;
            ng = min(ng + cnt, p.cand_cap);
          }
          float unused_worst;
          L.merge(d, link, cand_ok && admit, ef, scratch, topsize, unused_worst, ID_MASK);
          if (topsize > 0) lower = L.key_at(topsize - 1);                         // :320-321
        }
        __syncwarp();
        if (nv < 32) break;
      }
    }
    // reset the visited bits this query set
    __syncwarp();
    {
      const int nt = min(ntouched, p.touched_cap);
      for (int i = lane; i < nt; i += 32) vis[touched[i] >> 5] = 0u;
      if (ntouched > p.touched_cap) {  // log overflowed: wipe everything
        for (int64_t i = lane; i < p.visited_words; i += 32) vis[i] = 0u;
      }
    }
    if (overflow && lane == 0) atomicExch(p.overflow_flag, 1);
    __syncwarp();
    write_results<EPL>(g, p, q, L, lane, hops, nbrs, evals);
  }
}

// label list -> bitmap by internal id.  labels of the index are unique; the allowed list is
// first marked in a bitmap over label values, then gathered by internal id.
__global__ void mark_labels_kernel(const uint64_t *__restrict__ allowed, int64_t n, uint32_t *__restrict__ by_label,
                                   uint64_t max_label) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t l = allowed[i];
  if (l <= max_label) atomicOr(by_label + (l >> 5), 1u << (l & 31));
}
__global__ void gather_filter_kernel(const uint64_t *__restrict__ labels, int64_t n, const uint32_t *__restrict__ by_label,
                                     uint32_t *__restrict__ by_id) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    uint64_t l = labels[i];
    ok = (by_label[l >> 5] >> (l & 31)) & 1u;
  }
  unsigned m = __ballot_sync(FULL_MASK, ok);
  if ((threadIdx.x & 31) == 0 && (i < n)) by_id[i >> 5] = m;
}

// the reference layout's element size (4 + 4*maxM0 + code_row + 8) is not a multiple of 4 for odd
// code rows, so its fields are read bytewise
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// level-0 walk records from the reference-layout records: [links | neighbour codes]
__global__ void pack_rec0_kernel(const uint8_t *__restrict__ raw, int64_t n, int size_per_elem, int offset_data,
                                 int maxM0, int code_row, int code_off0, int rec0_bytes, uint8_t *__restrict__ rec0) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t node = t / maxM0;
  const int j = (int)(t - node * maxM0);
  if (node >= n) return;
  const uint8_t *src = raw + node * size_per_elem;
  const unsigned cnt = (unsigned)src[0] | ((unsigned)src[1] << 8);
  uint8_t *dst = rec0 + node * rec0_bytes;
  uint32_t link = EMPTY_LINK;
  if ((unsigned)j < cnt) link = ld_u32_unaligned(src + 4 + 4 * j);
  *reinterpret_cast<uint32_t *>(dst + 4 * j) = link;
  uint8_t *cdst = dst + code_off0 + (size_t)j * code_row;
  if (link != EMPTY_LINK) {
    const uint8_t *csrc = raw + (size_t)link * size_per_elem + offset_data;
    for (int b = 0; b < code_row; b++) cdst[b] = csrc[b];
  } else {
    for (int b = 0; b < code_row; b++) cdst[b] = 0;
  }
}

struct LaunchGeom {
  int warps, ctas_per_sm, smem_bytes;
  bool smem_table;
};

LaunchGeom pick_geometry(annb_index *h, size_t table_bytes_in, size_t extra_per_warp, int max_warps) {
  const size_t table_bytes = (table_bytes_in + 15) / 16 * 16 + extra_per_warp;
  int optin = 0, per_sm = 0;
  cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
  cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device);
  LaunchGeom best{4, 1, 0, false};
  if (table_bytes > (size_t)optin) {  // table does not fit: read it through L1/L2
    best.warps = std::min(8, max_warps);
    best.ctas_per_sm = 4;
    best.smem_bytes = (int)(best.warps * extra_per_warp);
    return best;
  }
  int best_total = 0;
  for (int W = 1; W <= max_warps; W++) {
    size_t need = W * table_bytes;
    if (need > (size_t)optin) break;
    int ctas = (int)((size_t)per_sm / (need + 1024));
    if (ctas < 1) continue;
    ctas = std::min(ctas, 32);
    int total = std::min(W * ctas, 32);  // beyond ~32 warps/SM the register file limits residency anyway
    if (total > best_total) {
      best_total = total;
      best.warps = W;
      best.ctas_per_sm = std::min(ctas, (32 + W - 1) / W);
      best.smem_bytes = (int)need;
      best.smem_table = true;
    }
  }
  if (h->opt_warps_per_cta > 0 && h->opt_warps_per_cta <= max_warps && h->opt_warps_per_cta * table_bytes <= (size_t)optin) {
    best.warps = (int)h->opt_warps_per_cta;
    best.smem_bytes = (int)(best.warps * table_bytes);
    best.ctas_per_sm = std::max(1, (int)((size_t)per_sm / (best.smem_bytes + 1024)));
    best.smem_table = true;
  }
  if (h->opt_ctas_per_sm > 0) best.ctas_per_sm = (int)h->opt_ctas_per_sm;
  return best;
}

template <int EPL, int CR, int CB>
int launch_walk(annb_index *h, const SearchParams &p_in, bool general) {
  SearchParams p = p_in;
  const size_t table_bytes = (size_t)h->M * h->Ks * sizeof(float);
  const bool chunked = h->gd.maxM0 > 32;
  const size_t extra = (size_t)(32 * EPL + (general ? SBAG : (chunked ? 0 : 33))) * sizeof(uint2);
  int max_warps = 32;
  LaunchGeom geo = pick_geometry(h, table_bytes, extra, max_warps);
  int table_stride = geo.smem_table ? (int)((table_bytes + 15) / 16 * 16) : 0;
  int threads = geo.warps * 32;
  int has_del = h->g.num_deleted > 0;
  int occ = geo.ctas_per_sm;
  bool fits = true;
  // Large EPL instantiations need > 64 registers/thread: a CTA of 32 warps may not fit the register
  // file.  ANNB_OCC asks the runtime; on 0 resident CTAs the geometry is re-picked with fewer warps.
#define ANNB_OCC(KERN)                                                                                                   \
  do {                                                                                                                   \
    for (;;) {                                                                                                           \
      if (geo.smem_bytes > 48 * 1024)                                                                                    \
        ANNB_CUDA(cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes));              \
      int o = 0;                                                                                                         \
      ANNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, KERN, threads, geo.smem_bytes));                       \
      if (o >= 1 || max_warps == 1) {                                                                                    \
        fits = o >= 1;                                                                                                   \
        occ = std::max(1, std::min(occ, o));                                                                             \
        break;                                                                                                           \
      }                                                                                                                  \
      max_warps = std::max(1, std::min(max_warps, geo.warps) / 2);                                                       \
      geo = pick_geometry(h, table_bytes, extra, max_warps);                                                             \
      table_stride = geo.smem_table ? (int)((table_bytes + 15) / 16 * 16) : 0;                                           \
      threads = geo.warps * 32;                                                                                          \
      occ = geo.ctas_per_sm;                                                                                             \
    }                                                                                                                    \
    if (!fits) ANNB_FAIL(ANNB_ELIMIT, "kernel does not fit on this device (registers/shared memory)");                   \
  } while (0)
  unsigned int *counter = p.work_counter;  // a caller running several launches concurrently passes its own slot
  int rc;
  if (!counter) {
    rc = annb_scratch(h, 4, 256, (void **)&counter);
    if (rc) return rc;
  }
  ANNB_CUDA(cudaMemsetAsync(counter, 0, 8, h->stream));
  p.work_counter = counter;
  p.overflow_flag = reinterpret_cast<int32_t *>(counter + 1);
  if (!general) {
#define ANNB_LAUNCH_FAST(KERN)                                                                                   \
  do {                                                                                                           \
    ANNB_OCC(KERN);                                                                                              \
    int blocks = (int)std::min<int64_t>((int64_t)h->sm_count * occ, (p.B + geo.warps - 1) / geo.warps);          \
    KERN<<<blocks, threads, geo.smem_bytes, h->stream>>>(h->gd, p, table_stride);                                \
  } while (0)
    if (chunked) {
      if (geo.smem_table) ANNB_LAUNCH_FAST((hnsw_walk_chunked<EPL, CR, CB, true>));
      else ANNB_LAUNCH_FAST((hnsw_walk_chunked<EPL, CR, CB, false>));
    } else {
      if (geo.smem_table) ANNB_LAUNCH_FAST((hnsw_walk_fast<EPL, CR, CB, true>));
      else ANNB_LAUNCH_FAST((hnsw_walk_fast<EPL, CR, CB, false>));
    }
#undef ANNB_LAUNCH_FAST
  } else {
    if (geo.smem_table) ANNB_OCC((hnsw_walk_general<EPL, CR, CB, true>));
    else ANNB_OCC((hnsw_walk_general<EPL, CR, CB, false>));
    int blocks = (int)std::min<int64_t>((int64_t)h->sm_count * occ, (p.B + geo.warps - 1) / geo.warps);
    int64_t slots = (int64_t)blocks * geo.warps;
    p.visited_words = (h->gd.n + 31) / 32;
    // per-warp scratch of the literal walk: visited log and candidate bag.  They start at sizes that cover ordinary
    // filters and grow (x4, up to the node count: neither can hold more) when a query outgrows them -- a very
    // selective filter or many deletions make the walk visit a large part of the graph, as the reference's does.
    const int64_t cap_max = std::max<int64_t>(1 << 15, h->gd.n + 64);
    int64_t tcap = 1 << 15, ccap = 1 << 14;
    for (;;) {
      p.touched_cap = (int)std::min<int64_t>(tcap, cap_max);
      p.cand_cap = (int)std::min<int64_t>(ccap, cap_max);
      // keep the per-warp scratch within ~8 GB: fewer resident warps when each needs a lot
      while (blocks > 1 && (double)blocks * geo.warps * ((double)p.touched_cap * 4 + (double)p.cand_cap * 8) > 8e9) blocks /= 2;
      slots = (int64_t)blocks * geo.warps;
      uint32_t *vis;
      static_assert(sizeof(unsigned int) == 4, "");
      // scratch slot 5 keeps the visited bitmaps; they are left all-zero by every query, so they are
      // cleared only when (re)allocated
      size_t vis_bytes = (size_t)slots * p.visited_words * 4;
      size_t before = h->scratch_cap[5];
      if ((rc = annb_scratch(h, 5, vis_bytes, (void **)&vis))) return rc;
      if (h->scratch_cap[5] != before) ANNB_CUDA(cudaMemsetAsync(vis, 0, h->scratch_cap[5], h->stream));
      p.visited = vis;
      if ((rc = annb_scratch(h, 6, (size_t)slots * p.touched_cap * 4, (void **)&p.touched))) return rc;
      if ((rc = annb_scratch(h, 7, (size_t)slots * p.cand_cap * 8, (void **)&p.cand))) return rc;
      ANNB_CUDA(cudaMemsetAsync(counter, 0, 8, h->stream));
      if (geo.smem_table)
        hnsw_walk_general<EPL, CR, CB, true><<<blocks, threads, geo.smem_bytes, h->stream>>>(h->gd, p, has_del, table_stride);
      else
        hnsw_walk_general<EPL, CR, CB, false><<<blocks, threads, geo.smem_bytes, h->stream>>>(h->gd, p, has_del, table_stride);
      h->launches++;
      ANNB_CUDA(cudaGetLastError());
      int32_t flag = 0;
      ANNB_CUDA(cudaMemcpyAsync(&flag, p.overflow_flag, 4, cudaMemcpyDeviceToHost, h->stream));
      ANNB_CUDA(cudaStreamSynchronize(h->stream));
      if (!flag) break;
      // visited bitmaps may be dirty after an overflow: wipe, then retry with more room
      ANNB_CUDA(cudaMemsetAsync(h->d_scratch[5], 0, h->scratch_cap[5], h->stream));
      if (p.touched_cap >= cap_max && p.cand_cap >= cap_max)
        ANNB_FAIL(ANNB_ELIMIT, "general walk scratch overflow (candidate bag %d / visited log %d entries per query)", p.cand_cap,
                  p.touched_cap);
      tcap *= 4;
      ccap *= 4;
    }
    return ANNB_OK;
  }
#undef ANNB_OCC
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

template <int EPL>
int dispatch_code(annb_index *h, const SearchParams &p, bool general) {
  const int cr = h->gd.code_row, cb = h->code_bytes;
  if (cb == 1) {
    switch (cr) {
      case 4: return launch_walk<EPL, 4, 1>(h, p, general);
      case 8: return launch_walk<EPL, 8, 1>(h, p, general);
      case 16: return launch_walk<EPL, 16, 1>(h, p, general);
      case 32: return launch_walk<EPL, 32, 1>(h, p, general);
      default: return launch_walk<EPL, 0, 1>(h, p, general);
    }
  } else {
    switch (cr) {
      case 8: return launch_walk<EPL, 8, 2>(h, p, general);
      case 16: return launch_walk<EPL, 16, 2>(h, p, general);
      default: return launch_walk<EPL, 0, 2>(h, p, general);
    }
  }
}

// ---- flagged walk launcher ---------------------------------------------------------------------------
template <int EPL, int CR, int CB>
int launch_flagged(annb_index *h, const SearchParams &p_in) {
  SearchParams p = p_in;
  const size_t table_bytes = (size_t)h->M * h->Ks * sizeof(float);
  const size_t extra = (size_t)(32 * EPL + 33) * sizeof(uint2);
  int max_warps = 32;
  LaunchGeom geo = pick_geometry(h, table_bytes, extra, max_warps);
  if (!geo.smem_table) return 1;  // not applicable: caller falls back to the bitmap walk
  auto kern = hnsw_walk_flagged<EPL, CR, CB>;
  int occ = 0, threads = 0, table_stride = (int)((table_bytes + 15) / 16 * 16);
  for (;;) {
    threads = geo.warps * 32;
    if (geo.smem_bytes > 48 * 1024) ANNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, geo.smem_bytes));
    ANNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, geo.smem_bytes));
    if (occ >= 1) break;
    if (max_warps == 1) return 1;
    max_warps = std::max(1, std::min(max_warps, geo.warps) / 2);
    geo = pick_geometry(h, table_bytes, extra, max_warps);
    if (!geo.smem_table) return 1;
  }
  occ = std::min(occ, geo.ctas_per_sm);
  unsigned int *counter = p.work_counter;
  if (!counter) {
    int rc = annb_scratch(h, 4, 256, (void **)&counter);
    if (rc) return rc;
  }
  ANNB_CUDA(cudaMemsetAsync(counter, 0, 8, h->stream));
  p.work_counter = counter;
  p.overflow_flag = reinterpret_cast<int32_t *>(counter + 1);
  const int blocks = (int)std::min<int64_t>((int64_t)h->sm_count * occ, (p.B + geo.warps - 1) / geo.warps);
  kern<<<blocks, threads, geo.smem_bytes, h->stream>>>(h->gd, p, h->g.num_deleted > 0 ? 1 : 0, table_stride);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

template <int EPL>
int dispatch_flagged(annb_index *h, const SearchParams &p) {
  const int cr = h->gd.code_row, cb = h->code_bytes;
  if (cb == 1) {
    switch (cr) {
      case 4: return launch_flagged<EPL, 4, 1>(h, p);
      case 8: return launch_flagged<EPL, 8, 1>(h, p);
      case 16: return launch_flagged<EPL, 16, 1>(h, p);
      case 32: return launch_flagged<EPL, 32, 1>(h, p);
      default: return 1;
    }
  }
  switch (cr) {
    case 8: return launch_flagged<EPL, 8, 2>(h, p);
    case 16: return launch_flagged<EPL, 16, 2>(h, p);
    default: return 1;
  }
}

}  // namespace

int launch_search(annb_index *h, const SearchParams &p, int mode) {
  if (p.B == 0) return ANNB_OK;
  if (p.B >= (int64_t)0xffffffffll) ANNB_FAIL(ANNB_ELIMIT, "at most 2^32-2 queries per call");
  if (mode == 1) {  // filter and/or deletions: two register lists in hnsw_walk4's mapping where that applies (walk_flagged4.cu)
    const int rc = launch_walk4f(h, p);
    if (rc != 1) return rc;
    if (!p.tables) ANNB_FAIL(ANNB_ESTATE, "internal: no tables for the flagged walk");
  }
  if (mode == 1 && h->gd.maxM0 <= 32 && h->gd.n < (1ll << 30)) {
    // list capacity the flagged walk needs: every candidate down to the ef-th admitted one stays listed
    const double s = std::min(1.0, std::max(1e-4, (double)p.selectivity));
    // until the ef-th admitted node is found the list holds every evaluated node: about ef/s of them
    // (binomial spread ~ sqrt(ef/s)); 1.3x + 48 leaves > 4 sigma of head-room at the usual sizes
    const double need = (double)p.ef / s * 1.3 + 48.0;
    int epl = 0;
    for (int e : {2, 4, 8, 16})
      if (epl == 0 && e * 32 >= need && e * 32 >= p.ef) epl = e;
    if (h->opt_flagged_epl > 0 && h->opt_flagged_epl * 32 >= p.ef) epl = (int)h->opt_flagged_epl;
    int rc = 1;
    if (epl == 2) rc = dispatch_flagged<2>(h, p);
    else if (epl == 4) rc = dispatch_flagged<4>(h, p);
    else if (epl == 8) rc = dispatch_flagged<8>(h, p);
    else if (epl == 16) rc = dispatch_flagged<16>(h, p);
    if (rc != 1) return rc;  // launched (or failed for real); 1 = not applicable -> bitmap walk
  }
  const bool general = mode != 0;
  if (!general && walk4_applicable(h)) {  // the plain search: K1 fused into the walk (walk_fused.cu)
    const int rc = launch_walk4(h, p);
    if (rc != 1) return rc;
  }
  if (!p.tables) ANNB_FAIL(ANNB_ESTATE, "internal: no tables for the walk");
  const int epl = (p.ef + 31) / 32;
  if (epl <= 2) return dispatch_code<2>(h, p, general);
  if (epl <= 4) return dispatch_code<4>(h, p, general);
  if (epl <= 8) return dispatch_code<8>(h, p, general);
  if (epl <= 16) return dispatch_code<16>(h, p, general);
  ANNB_FAIL(ANNB_ELIMIT, "ef=%d exceeds ANNB_MAX_EF=%d", p.ef, ANNB_MAX_EF);
}

int launch_pack_rec0(annb_index *h, const uint8_t *d_level0_raw, int64_t n) {
  if (n == 0) return ANNB_OK;
  const int64_t total = n * h->gd.maxM0;
  pack_rec0_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(
      d_level0_raw, n, (int)h->g.size_per_elem, (int)h->g.offset_data, h->gd.maxM0, h->gd.code_row, h->gd.code_off0,
      h->gd.rec0_bytes,
      h->d_rec0);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

// patch of the device graph: staged[i] (one packed record) -> dst[ids[i]]
__global__ void scatter_records_kernel(const uint8_t *__restrict__ staged, const uint32_t *__restrict__ ids, int64_t cnt, int words,
                                       uint8_t *__restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / words;
  if (i >= cnt) return;
  const int w = (int)(t - i * words);
  reinterpret_cast<uint32_t *>(dst + (size_t)ids[i] * words * 4)[w] = reinterpret_cast<const uint32_t *>(staged + (size_t)i * words * 4)[w];
}

int launch_scatter_records(annb_index *h, const uint8_t *d_staged, const uint32_t *d_ids, int64_t cnt, int rec_bytes, uint8_t *d_dst) {
  if (cnt == 0) return ANNB_OK;
  const int words = rec_bytes / 4;
  const int64_t total = cnt * words;
  scatter_records_kernel<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(d_staged, d_ids, cnt, words, d_dst);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

int launch_filter_bitmap(annb_index *h, const uint64_t *d_filter_labels, int64_t n_filter, uint32_t *d_by_label,
                         uint32_t *d_by_id) {
  const int64_t n = h->gd.n;
  const size_t words = (size_t)(h->max_label >> 5) + 1;
  ANNB_CUDA(cudaMemsetAsync(d_by_label, 0, words * 4, h->stream));
  if (n_filter > 0)
    mark_labels_kernel<<<(unsigned)((n_filter + 255) / 256), 256, 0, h->stream>>>(d_filter_labels, n_filter, d_by_label,
                                                                               h->max_label);
  const int64_t padded = (n + 31) / 32 * 32;
  gather_filter_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, h->stream>>>(h->d_labels, n, d_by_label, d_by_id);
  h->launches += 2;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}
