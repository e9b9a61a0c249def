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
};
