// walk_fused.cu -- K1+K3 in one kernel: the warp that walks a query first builds that query's ADC table
// straight into its own shared-memory slot, then runs the HNSW walk over PQ codes out of registers.
//
// Reference semantics (unchanged from hnsw_search.cu, where the equivalence argument lives):
//   table : batch_precompute_adc_table[_ip]  bindings/pq_bindings.pyx:149-274 (+ pq.py:316-322 epilogue)
//           T[m][c] = sum_j (cb[m][c][j] - q[m*ds+j])^2, j sequential, sub / mul / add rounded separately
//   walk  : searchKnn  include/hnswlib/hnswalg.h:1237-1295 = greedy descent (:1248-1274) +
//           searchBaseLayerST (:243-329) with PQLookup (include/hnswlib/space_pq.h:16-37)
//
// What is different from hnsw_walk_fast (round 1), and why -- that kernel was issue-bound at ~305 warp
// instructions per hop:
//   * compile-time M and Ks = 256, u8 codes: a table lookup is byte-extract + LDS with an immediate
//     row offset + FADD; no run-time m*Ks+code / node*rec_bytes integer chains.
//   * the ef-bounded result list lives ONLY in registers, blocked (lane l holds positions l*EPL..),
//     keys as order-preserving u32 images of the fp32 distances so that warp minima are one REDUX.
//     No shared-memory mirror, no binary search, no rank loops, no __syncwarp in the hop.
//   * candidates of a hop are consumed smallest-first: the first one that is not already listed is, if
//     it beats the nearest unexpanded list entry, exactly the next node to expand -- so the next record
//     is requested at that moment (exact, never a wrong guess) and its DRAM round trip overlaps the
//     insertions.  Consuming in ascending order also lets lowerBound fall as fast as it can: the loop
//     stops at the first candidate that no longer beats it (the hop-start admission test of the
//     single-list walk admits a superset whose tail falls off the list; the final list is identical).
//   * one insertion = two shuffles (carry from the lane below) + per-slot compare/select; equal keys
//     keep arrival order (new after old, lower lane first), the tie rule of the single-list model.
//   * shared memory per query = the table and nothing else: 28 queries per SM at M=8 (was 25).
//   * K1 fused: the table never exists in HBM.  The codebook is read from L2 through a transposed copy
//     ([m][j/V][c][V], V = 4 or 2 floats) so that a warp's codeword loads are fully coalesced.
//     The literal `tables=` call form (the reference's dtables argument) is served by the same kernel:
//     the table is then staged by one TMA bulk copy (cp.async.bulk + mbarrier) per query.
//
// Results, hop and neighbour counters are those of hnsw_walk_fast and of oracle.single_list_walk on
// every input (same list after every hop, same next node); differences from the reference itself are
// confined to exact fp32 ties, as before.
#include <math_constants.h>

#include <algorithm>

#include "annb_internal.h"
#include "tma_utils.cuh"
#include "walk4_common.cuh"

namespace {

// Launch geometry.  Shared memory per query = its table (M KB); the TMA form adds one 8-byte mbarrier per warp.
// An SM has 228 KB of shared memory and every resident CTA costs 1 KB on top of what it asks for, so at M=8
// 4 CTAs x 7 warps x 8 KB + 4 KB fills it to the byte: 28 queries per SM in CTAs small enough that the tail of one
// launch frees SMs piecewise for the next (the two lanes of annb_search_submit overlap that way).  The TMA form
// (8 bytes more per warp) runs 3 CTAs x 9 warps = 27.  Wider lists need more registers: fewer warps.
constexpr int W4_SMEM_SM = 233472;  // 228 KB per SM
constexpr int w4_cta_warps(int M, int EPL, bool fused) {
  return M == 8 ? (EPL <= 4 ? (fused ? 7 : 9) : (EPL == 8 ? 5 : 7)) : 7;
}
constexpr int w4_ctas(int M, int EPL, bool fused) {
  return M == 8 ? (EPL <= 4 ? (fused ? 4 : 3) : (EPL == 8 ? 4 : 2)) : (M == 16 ? 2 : 1);
}
constexpr int w4_smem(int M, int EPL, bool fused) { return w4_cta_warps(M, EPL, fused) * (M * 1024 + (fused ? 0 : 8)); }
constexpr bool w4_fits(int M, int EPL, bool fused) {
  return w4_ctas(M, EPL, fused) * (w4_smem(M, EPL, fused) + 1024) <= W4_SMEM_SM;
}
constexpr int w4_max_threads(int M, int EPL) {
  return 32 * (w4_cta_warps(M, EPL, true) > w4_cta_warps(M, EPL, false) ? w4_cta_warps(M, EPL, true) : w4_cta_warps(M, EPL, false));
}
constexpr int w4_min_ctas(int M, int EPL) {  // register cap = 64K / (max threads x this): must allow either residency
  const int tf = w4_ctas(M, EPL, true) * w4_cta_warps(M, EPL, true) * 32, tt = w4_ctas(M, EPL, false) * w4_cta_warps(M, EPL, false) * 32;
  return (tf > tt ? tf : tt) / w4_max_threads(M, EPL);
}

// =====================================================================================================
template <int M, int EPL>
__global__ void __launch_bounds__(w4_max_threads(M, EPL), w4_min_ctas(M, EPL)) hnsw_walk4(const GraphDev g, const SearchParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  // warp index through a warp reduction: the compiler then knows it is warp-uniform, keeps the table base in
  // a uniform register and folds it into the LDS address ([R + UR + imm]) -- one IADD less per lookup
  const int warp = (int)__reduce_min_sync(FULL_MASK, threadIdx.x >> 5);
  const int nwarps = blockDim.x >> 5;
  constexpr int TS = M * 256;
  constexpr int CAP = 32 * EPL;
  float *T = reinterpret_cast<float *>(smem_raw) + (size_t)warp * TS;
  uint64_t *tbar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)nwarps * TS * 4) + warp;
  const bool fused = p.queries != nullptr;
  uint32_t tphase = 0;
  if (!fused) {
    if (lane == 0) mbar_init(tbar, 1);
    __syncwarp();
  }
  const int ef = p.ef;
  const int k = p.k;
  const bool lane0 = lane == 0;
  const bool stats = p.out_stats != nullptr;
  // list positions ef-1 (lowerBound once full) and ef (the slot an insertion into a full list spills to)
  const int wl = (ef - 1) / EPL, ws = (ef - 1) % EPL;
  // per-lane record addressing
  const bool has_slot0 = lane < g.maxM0;
  const uint8_t *link_base0 = g.rec0 + 4 * lane;
  {  // keep the per-lane base as ONE 64-bit register pair (no re-derivation from the constant bank per load)
    uint64_t t = reinterpret_cast<uint64_t>(link_base0);
    asm volatile("" : "+l"(t));
    link_base0 = reinterpret_cast<const uint8_t *>(t);
  }
  const unsigned code_delta0 = (unsigned)(g.code_off0 + lane * (M - 4));  // from a lane's link to the same lane's code row
  const uint32_t rec0_bytes = (uint32_t)g.rec0_bytes;
  const bool has_slotu = lane < g.maxM;

  for (;;) {
    // next query: ONE lane bumps the work counter and the value is spread with a warp REDUCTION, not a shuffle:
    // REDUX delivers it in a uniform register, so ptxas can prove that the loop exit below is warp-uniform.
    // With `__shfl_sync(qi, 0)` it cannot, and then guards every warp-level operation of the whole loop body with
    // a BRA.DIV convergence check (31 sites, ~25 extra warp instructions per hop).
    unsigned qi = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %2, 0;\n"
        "@p atom.global.add.u32 %0, [%1], 1;\n"
        "}\n"
        : "+r"(qi)
        : "l"(p.work_counter), "r"(lane)
        : "memory");
    const int64_t q = __reduce_max_sync(FULL_MASK, qi);
    if (q >= p.B) break;
    __syncwarp();  // every lane is done with the previous query's table
    if (fused) {
      const float *qg = p.queries + q * (int64_t)(M * p.ds);
      if (p.cb_vec == 4) build_table<M, 4>(T, qg, p.cbt, p.ds, p.is_ip, p.bias, lane);
      else build_table<M, 2>(T, qg, p.cbt, p.ds, p.is_ip, p.bias, lane);
    } else {
      if (lane0) {
        mbar_expect_tx(tbar, (uint32_t)TS * 4u);
        bulk_g2s(T, p.tables + q * TS, (uint32_t)TS * 4u, tbar);
      }
      mbar_wait(tbar, tphase);
      tphase ^= 1u;
    }
    if (p.dump_tables) {  // debug export of the in-shared-memory table (parity test of the fused build)
      float *o = p.dump_tables + q * TS;
      for (int i = lane; i < TS; i += 32) o[i] = T[i];
    }

    // ---- greedy descent maxlevel..1 (hnswalg.h:1245-1274), walk4_common.cuh ----
    int hops = 0, nbrs = 0, evals = 1;
    uint32_t cur_uk, rec;
    descend4<M>(g, T, lane, has_slotu, hops, nbrs, evals, cur_uk, rec);
    evals += 1;  // searchBaseLayerST re-scores the entry (:255)

    // ---- level 0: searchBaseLayerST as a single sorted list in registers ----
    uint32_t K[EPL], V[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      K[e] = KEY_MAX;
      V[e] = 0xffffffffu;
    }
    if (lane0) {
      K[0] = cur_uk;
      V[0] = rec | EXP_BIT;  // the entry node is hop 0
    }
    int size = 1;
    uint32_t worst = (ef == 1) ? cur_uk : KEY_MAX;  // lowerBound (:306); "infinite" while the list is not full

    Rec<M> recA, recB;
    auto load_rec = [&](Rec<M> &r, uint32_t node) {
      const uint8_t *lp = link_base0 + (size_t)node * rec0_bytes;
      r.link = EMPTY_LINK;
#pragma unroll
      for (int i = 0; i < M / 4; i++) r.cw[i] = 0u;
      if (has_slot0) {
        r.link = __ldg(reinterpret_cast<const uint32_t *>(lp));
        load_codes<M>(r.cw, lp + code_delta0);
      }
    };
    load_rec(recA, rec);

    // One hop: `cur` holds the record of the node being expanded, `nxt` receives the record of the next one.
    // The two record register sets alternate roles (no copy, so nothing waits on a load before it is needed).
    auto hop = [&](Rec<M> &cur, Rec<M> &nxt) -> bool {
      hops++;
      // ---- nearest not-yet-expanded list entry == candidate_set.top() (:268).  It is the next node unless this
      // hop finds something closer, so its record is pulled into L2 FIRST: the DRAM round trip then overlaps the
      // whole hop, and a wrong guess still leaves the record in L2 for the hop that does expand it ----
      uint32_t lm = KEY_MAX;
#pragma unroll
      for (int e = 0; e < EPL; e++) lm = min(lm, K[e] | (uint32_t)((int32_t)V[e] >> 31));
      const uint32_t e2key = __reduce_min_sync(FULL_MASK, lm);
      int e2src = 0;
      uint32_t e2id = 0;
      if (e2key != KEY_MAX) {
        e2src = __ffs(__ballot_sync(FULL_MASK, lm == e2key)) - 1;
        uint32_t myid = 0;
#pragma unroll
        for (int e = EPL - 1; e >= 0; e--)
          if ((K[e] | (uint32_t)((int32_t)V[e] >> 31)) == e2key) myid = V[e] & IDM;  // lowest matching slot wins
        e2id = __shfl_sync(FULL_MASK, myid, e2src);
        // (an L2 prefetch, not a register load: a second load into the same registers would have to wait for
        // this one to land whenever a new candidate wins)
        if (p.prefetch & 4) {         // bulk form: the TMA unit fetches the whole record on one lane's request
          if (lane0) bulk_prefetch_l2(g.rec0 + (size_t)e2id * rec0_bytes, rec0_bytes);
        } else if ((p.prefetch & 1) && lane * 128u < rec0_bytes) {
          prefetch_l2(g.rec0 + (size_t)e2id * rec0_bytes + lane * 128u);
        }
      }
      // expand that entry: flag it (the lowest matching slot of lane e2src)
      auto take_e2 = [&]() {
        if (lane == e2src) {
          bool done = false;
#pragma unroll
          for (int e = 0; e < EPL; e++) {
            const bool hit = !done && (K[e] | (uint32_t)((int32_t)V[e] >> 31)) == e2key;
            if (hit) V[e] |= EXP_BIT;
            done |= hit;
          }
        }
      };

      // ---- score the neighbour list: one lane = one neighbour, m sequential ----
      const bool valid = cur.link != EMPTY_LINK;
      const uint32_t uk = f2u(pq_score<M>(T, cur.cw));
      uint32_t mykey = (valid && uk < worst) ? uk : KEY_MAX;  // admission (:306) against the hop-start lowerBound
      if (stats) {
        const int nv = __popc(__ballot_sync(FULL_MASK, valid));
        nbrs += nv;
        evals += nv;
      }
      if ((p.prefetch & 2) && mykey < e2key) {
        // closer than everything still unexpanded: (one of) the next node(s) -- warm its record in L2 now
        const uint8_t *rp = g.rec0 + (size_t)cur.link * rec0_bytes;
        for (uint32_t o = 0; o < rec0_bytes; o += 128) prefetch_l2(rp + o);
      }

      // re-encounter of a listed node?  (an id can only be listed under this very key)
      auto listed = [&](uint32_t id) -> bool {
        bool dup = false;
#pragma unroll
        for (int e = 0; e < EPL; e++) dup |= (V[e] & IDM) == id;
        return __any_sync(FULL_MASK, dup);
      };

      // ---- phase A: the smallest NEW candidate (if any) against the nearest unexpanded entry: the next node is
      // known exactly, before anything is merged ----
      bool have_new = false;
      uint32_t mn, id = 0;
      for (;;) {
        mn = __reduce_min_sync(FULL_MASK, mykey);
        if (mn == KEY_MAX) break;  // none (left) that beats lowerBound
        const int src = __ffs(__ballot_sync(FULL_MASK, mykey == mn)) - 1;  // lower lane first among equals
        id = __shfl_sync(FULL_MASK, cur.link, src);
        if (lane == src) mykey = KEY_MAX;
        if (!listed(id)) {
          have_new = true;
          break;
        }
      }
      if (have_new && mn < e2key) {
        load_rec(nxt, id);  // replaces the guess; its DRAM round trip overlaps the insertions below
        list_insert<EPL>(K, V, mn, id | EXP_BIT, lane0);
        size++;
      } else {
        if (e2key == KEY_MAX) return false;  // nothing unexpanded and nothing new: candidate_set exhausted (:266)
        load_rec(nxt, e2id);
        take_e2();
        if (have_new) {
          list_insert<EPL>(K, V, mn, id, lane0);
          size++;
        }
      }
      // ---- phase B: the other admitted candidates in lane order (equal keys keep arrival order).  lowerBound
      // is NOT refreshed per insertion: a candidate that the falling bound would have rejected lands at a
      // position >= ef and drops out below, exactly like the hop-start admission rule of the single-list walk ----
      if (have_new) {
        // (A/B, round 2: checking the re-encounters of a hop four at a time ahead of the insertions, with the next
        // candidate's key/id fetched before the current insertion, was 13 % SLOWER -- 30.4 vs 34.8 M QPS on one box:
        // the padded groups cost more issue slots than the overlapped latencies give back at ~3.4 candidates per hop)
        unsigned live = __ballot_sync(FULL_MASK, mykey != KEY_MAX);
        while (live) {
          const int src = __ffs(live) - 1;
          live &= live - 1;
          const uint32_t ck = __shfl_sync(FULL_MASK, mykey, src);
          const uint32_t cid = __shfl_sync(FULL_MASK, cur.link, src);
          if (listed(cid)) continue;
          list_insert<EPL>(K, V, ck, cid, lane0);
          size++;
        }
        if (size >= ef) {
          size = ef;
          worst = list_key_at<EPL>(K, wl, ws);  // lowerBound = top_candidates.top() (:320-321)
          if (ef < CAP) {                        // what spilled past position ef-1 is gone
#pragma unroll
            for (int e = 0; e < EPL; e++)
              if (lane * EPL + e >= ef) {
                K[e] = KEY_MAX;
                V[e] = 0xffffffffu;
              }
          }
        }
      }
      return true;
    };
    for (;;) {
      if (!hop(recA, recB)) break;
      if (!hop(recB, recA)) break;
    }

    // ---- results: the first k list entries, ascending (dist, label) (hnsw_bindings.cpp:346-351) ----
    const int found = min(size, k);
    bool tie = false;
    const uint32_t k_next_lane = __shfl_down_sync(FULL_MASK, K[0], 1);
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int pos = lane * EPL + e;
      if (pos < k) {
        const bool have = pos < size;
        p.out_dists[q * k + pos] = have ? u2f(K[e]) : CUDART_INF_F;
        p.out_labels[q * k + pos] = have ? (p.out_internal ? (uint64_t)(V[e] & IDM) : __ldg(g.labels + (V[e] & IDM))) : (uint64_t)UINT64_MAX;
      }
      const uint32_t nk = (e + 1 < EPL) ? K[e + 1 < EPL ? e + 1 : e] : k_next_lane;
      tie |= (pos + 1 < found) && (nk == K[e]) && (e + 1 < EPL || lane < 31);
    }
    if (__any_sync(FULL_MASK, tie)) {
      // rows must be ascending by (dist, label); equal distances are rare, fix them up serially
      __syncwarp();
      if (lane0) {
        for (int i = 1; i < found; i++) {
          const float d = p.out_dists[q * k + i];
          const uint64_t l = p.out_labels[q * k + i];
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
    if (lane0) {
      p.out_found[q] = found;
      if (stats) {
        p.out_stats[q * 3 + 0] = hops;
        p.out_stats[q * 3 + 1] = nbrs;
        p.out_stats[q * 3 + 2] = evals;
      }
    }
  }
}

template <int M, int EPL>
int launch_walk4_t(annb_index *h, const SearchParams &p_in) {
  SearchParams p = p_in;
  p.prefetch = (int)h->opt_prefetch;
  static_assert(w4_fits(M, EPL, true) && w4_fits(M, EPL, false), "walk4 geometry exceeds the SM's shared memory");
  auto kern = hnsw_walk4<M, EPL>;
  const bool fused = p.queries != nullptr;
  int warps = fused ? w4_cta_warps(M, EPL, true) : w4_cta_warps(M, EPL, false);
  int ctas = fused ? w4_ctas(M, EPL, true) : w4_ctas(M, EPL, false);
  if (h->opt_warps_per_cta > 0) warps = (int)std::min<int64_t>(h->opt_warps_per_cta, warps);
  if (h->opt_ctas_per_sm > 0) ctas = (int)std::min<int64_t>(h->opt_ctas_per_sm, ctas);
  // small batches: spread the queries over all SMs instead of filling a few of them
  const int64_t per_sm = (p.B + h->sm_count - 1) / h->sm_count;
  if (per_sm < (int64_t)warps * ctas) {
    ctas = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (per_sm + warps - 1) / warps));
    if (ctas == 1) warps = (int)std::max<int64_t>(1, std::min<int64_t>(warps, per_sm));
  }
  const int smem = warps * (M * 1024 + (fused ? 0 : 8));
  ANNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 std::max(w4_smem(M, EPL, true), w4_smem(M, EPL, false))));
  unsigned int *counter = p.work_counter;
  if (!counter) {
    int rc = annb_scratch(h, 4, 256, (void **)&counter);
    if (rc) return rc;
  }
  ANNB_CUDA(cudaMemsetAsync(counter, 0, 8, h->stream));
  p.work_counter = counter;
  p.overflow_flag = reinterpret_cast<int32_t *>(counter + 1);
  const int blocks = (int)std::min<int64_t>((int64_t)h->sm_count * ctas, (p.B + warps - 1) / warps);
  kern<<<blocks, warps * 32, smem, h->stream>>>(h->gd, p);
  h->launches++;
  ANNB_CUDA(cudaGetLastError());
  return ANNB_OK;
}

template <int M>
int launch_walk4_m(annb_index *h, const SearchParams &p) {
  const int epl = (p.ef + 31) / 32;
  if (epl <= 2) return launch_walk4_t<M, 2>(h, p);
  if (epl <= 4) return launch_walk4_t<M, 4>(h, p);
  if (epl <= 8) return launch_walk4_t<M, 8>(h, p);
  return launch_walk4_t<M, 16>(h, p);
}

}  // namespace

// Can the plain (no filter, no deletions) search of this index run on hnsw_walk4 at all?
bool walk4_applicable(const annb_index *h) {
  if (h->opt_walk_kernel == 1) return false;  // A/B switch: force the round-1 kernels
  const GraphDev &d = h->gd;
  return h->Ks == 256 && h->code_bytes == 1 && (h->M == 8 || h->M == 16 || h->M == 32) && d.maxM0 <= 32 && d.maxM <= 32 &&
         d.n < 0x7fffffffll;
}

// ... and may it build the tables itself from the queries (fused K1)?  Needs a vectorisable subvector
// length and a staging layout that is consumed before it is overwritten (see build_table).
bool walk4_can_fuse(const annb_index *h) {
  if (h->opt_walk_kernel == 2) return false;  // A/B switch: hnsw_walk4 on materialised tables
  if (!walk4_applicable(h) || !h->d_codebook_t || h->cb_vec == 0) return false;
  const int M = h->M, ds = h->ds, D = M * ds;
  const int R = (D + 255) / 256;
  if (R > M) return false;
  for (int t = 0; t + 1 < R; t++)
    if ((t + 1) * 256 > (M - R + t + 1) * ds) return false;
  return true;
}

int launch_walk4(annb_index *h, const SearchParams &p) {
  if (p.B == 0) return ANNB_OK;
  if (!p.queries && (reinterpret_cast<uintptr_t>(p.tables) & 15)) return 1;  // TMA needs 16-byte alignment
  if (p.ef > ANNB_MAX_EF) ANNB_FAIL(ANNB_ELIMIT, "ef=%d exceeds ANNB_MAX_EF=%d", p.ef, ANNB_MAX_EF);
  switch (h->M) {
    case 8: return launch_walk4_m<8>(h, p);
    case 16: return launch_walk4_m<16>(h, p);
    case 32: return launch_walk4_m<32>(h, p);
  }
  return 1;
}
