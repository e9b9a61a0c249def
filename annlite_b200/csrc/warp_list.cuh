// warp_list.cuh -- a sorted (ascending) bounded list of (fp32 key, u32 value) pairs held in the
// registers of one warp: EPL entries per lane, striped so that list position p lives in
// register row p/32 of lane p%32.  This is the GPU stand-in for the reference's
// `top_candidates` max-heap of at most ef entries (hnswalg.h:250, :306-322): the heap's top
// (lowerBound) is simply the key at position cap-1, and insertion keeps arrival order among
// equal keys (new entries go after existing equals), which is what makes ties deterministic.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define FULL_MASK 0xffffffffu
#define LIST_EMPTY_VAL 0xffffffffu

template <int EPL>
struct WarpList {
  float k[EPL];
  uint32_t v[EPL];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      k[e] = CUDART_INF_F;
      v[e] = LIST_EMPTY_VAL;
    }
  }

  // key at list position pos (warp-uniform pos), broadcast to all lanes
  __device__ __forceinline__ float key_at(int pos) const {
    float r = CUDART_INF_F;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      float t = __shfl_sync(FULL_MASK, k[e], pos & 31);
      if ((pos >> 5) == e) r = t;
    }
    return r;
  }
  __device__ __forceinline__ uint32_t val_at(int pos) const {
    uint32_t r = LIST_EMPTY_VAL;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      uint32_t t = __shfl_sync(FULL_MASK, v[e], pos & 31);
      if ((pos >> 5) == e) r = t;
    }
    return r;
  }

  // number of occupied positions (keys < +inf); list is dense from position 0
  __device__ __forceinline__ int size() const {
    int n = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) n += __popc(__ballot_sync(FULL_MASK, v[e] != LIST_EMPTY_VAL));
    return n;
  }

  // true if some entry's (value & mask) equals id
  __device__ __forceinline__ bool contains(uint32_t id, uint32_t mask) const {
    bool hit = false;
#pragma unroll
    for (int e = 0; e < EPL; e++) hit |= ((v[e] & mask) == id) && (v[e] != LIST_EMPTY_VAL);
    return __any_sync(FULL_MASK, hit);
  }

  // Insert (d, val) keeping ascending order, after any equal keys; entries pushed beyond
  // position cap-1 fall off (cap <= 32*EPL, warp-uniform).  All lanes pass the same (d, val).
  // Returns the insert position, or -1 if it fell beyond cap.
  __device__ __forceinline__ int insert(float d, uint32_t val, int cap) {
    const int lane = threadIdx.x & 31;
    int pos = 0;
#pragma unroll
    for (int e = 0; e < EPL; e++) pos += __popc(__ballot_sync(FULL_MASK, k[e] <= d));
    if (pos >= cap) return -1;
#pragma unroll
    for (int e = EPL - 1; e >= 0; e--) {
      if ((e + 1) * 32 <= pos) break;  // rows entirely before the insert point are untouched
      float upk = __shfl_up_sync(FULL_MASK, k[e], 1);
      uint32_t upv = __shfl_up_sync(FULL_MASK, v[e], 1);
      float ck = 0.f;
      uint32_t cv = 0;
      if (e > 0) {
        ck = __shfl_sync(FULL_MASK, k[e - 1], 31);
        cv = __shfl_sync(FULL_MASK, v[e - 1], 31);
      }
      if (lane == 0) {
        upk = ck;
        upv = cv;
      }
      const int mypos = e * 32 + lane;
      if (mypos > pos) {
        k[e] = upk;
        v[e] = upv;
      } else if (mypos == pos) {
        k[e] = d;
        v[e] = val;
      }
      if (mypos >= cap) {
        k[e] = CUDART_INF_F;
        v[e] = LIST_EMPTY_VAL;
      }
    }
    return pos;
  }

  // Merge up to 32 offered candidates (one per lane, `take` lanes offer (d, val)) into the list in
  // ONE pass instead of 32 serial inserts.  Every candidate's final position is computed by counting
  // (list entries <= d) + (live candidates ordered before it by (d, lane)); every list entry moves
  // right by the number of live candidates with a strictly smaller key; the merged sequence is
  // scattered through a per-warp shared-memory scratch (32*EPL uint2) and read back.  Ordering is
  // identical to inserting the candidates one at a time in lane order (equal keys keep arrival
  // order, the tail beyond `cap` falls off).  A candidate whose (key, id) already sits in the list is
  // dropped (ids are unique per node, so an equal id implies an equal key: only equal keys are
  // id-checked).  `size` and `worst` (key at cap-1, +inf while not full) are warp-uniform state.
  __device__ __forceinline__ void merge(float d, uint32_t val, bool take, int cap, uint2 *scratch, int &size,
                                        float &worst, uint32_t id_mask) {
    const int lane = threadIdx.x & 31;
    const unsigned offered = __ballot_sync(FULL_MASK, take);
    if (!offered) return;
    int shift[EPL];
#pragma unroll
    for (int e = 0; e < EPL; e++) shift[e] = 0;
    int rank = 0, base = 0;
    unsigned live = offered, m = offered;
    while (m) {
      const int c = __ffs(m) - 1;
      m &= m - 1;
      const float dc = __shfl_sync(FULL_MASK, d, c);
      int cnt = 0;
      bool eq = false;
#pragma unroll
      for (int e = 0; e < EPL; e++) {
        cnt += __popc(__ballot_sync(FULL_MASK, k[e] <= dc));
        eq |= (k[e] == dc);
      }
      if (__any_sync(FULL_MASK, eq)) {
        const uint32_t idc = __shfl_sync(FULL_MASK, val, c) & id_mask;
        bool dup = false;
#pragma unroll
        for (int e = 0; e < EPL; e++) dup |= (k[e] == dc) && ((v[e] & id_mask) == idc) && (v[e] != LIST_EMPTY_VAL);
        if (__any_sync(FULL_MASK, dup)) {
          live &= ~(1u << c);
          continue;
        }
      }
#pragma unroll
      for (int e = 0; e < EPL; e++) shift[e] += (k[e] > dc) ? 1 : 0;
      if (lane == c) base = cnt;
      rank += ((dc < d) || (dc == d && c < lane)) ? 1 : 0;
    }
    if (!live) return;
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int np = e * 32 + lane + shift[e];
      if (v[e] != LIST_EMPTY_VAL && np < cap) scratch[np] = make_uint2(__float_as_uint(k[e]), v[e]);
    }
    if ((live >> lane) & 1u) {
      const int np = base + rank;
      if (np < cap) scratch[np] = make_uint2(__float_as_uint(d), val);
    }
    size = min(size + __popc(live), cap);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < EPL; e++) {
      const int pos = e * 32 + lane;
      if (pos < size) {
        const uint2 t = scratch[pos];
        k[e] = __uint_as_float(t.x);
        v[e] = t.y;
      } else {
        k[e] = CUDART_INF_F;
        v[e] = LIST_EMPTY_VAL;
      }
    }
    worst = (size == cap) ? __uint_as_float(scratch[cap - 1].x) : CUDART_INF_F;
    __syncwarp();
  }
};
