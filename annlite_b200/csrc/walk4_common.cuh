// walk4_common.cuh -- pieces shared by hnsw_walk4 (walk_fused.cu) and hnsw_walk4f (walk_flagged4.cu): the
// order-preserving key image, a level-0 record in registers, the PQ lookup with compile-time M and Ks = 256, and the
// register-resident sorted list (blocked: lane l holds positions l*EPL .. l*EPL+EPL-1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef FULL_MASK
#define FULL_MASK 0xffffffffu
#endif

namespace {

constexpr uint32_t EMPTY_LINK = 0xffffffffu;
constexpr uint32_t EXP_BIT = 0x80000000u;   // list value: node already expanded (empty slots carry it too)
constexpr uint32_t IDM = 0x7fffffffu;
constexpr uint32_t KEY_MAX = 0xffffffffu;   // above the image of every finite distance and of +inf

// order-preserving map fp32 -> u32 (a < b  <=>  f2u(a) < f2u(b) for non-NaN, no -0.0: the ADC sum starts
// at +0.f and x + (-0) never yields -0)
__device__ __forceinline__ uint32_t f2u(float f) {
  const uint32_t b = __float_as_uint(f);
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float u2f(uint32_t u) {
  return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u));
}

// ---- one level-0 record in registers: lane j holds link j and the M code bytes of neighbour j ----
template <int M>
struct Rec {
  uint32_t link;
  uint32_t cw[M / 4];
};

template <int M>
__device__ __forceinline__ void load_codes(uint32_t (&cw)[M / 4], const uint8_t *p) {
  if (M == 4) {
    cw[0] = __ldg(reinterpret_cast<const uint32_t *>(p));
  } else if (M == 8) {
    const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
    cw[0] = t.x;
    cw[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < M / 16; i++) {
      const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p) + i);
      cw[4 * i + 0] = t.x;
      cw[4 * i + 1] = t.y;
      cw[4 * i + 2] = t.z;
      cw[4 * i + 3] = t.w;
    }
  }
}

// PQLookup (space_pq.h:30-35): strictly sequential fp32 sum over the M subquantisers, from 0.f
template <int M>
__device__ __forceinline__ float pq_score(const float *T, const uint32_t (&cw)[M / 4]) {
  float r = 0.f;
#pragma unroll
  for (int m = 0; m < M; m++) {
    const uint32_t code = (cw[m >> 2] >> (8 * (m & 3))) & 0xffu;
    r = __fadd_rn(r, T[m * 256 + code]);
  }
  return r;
}

// ---- the list: EPL entries per lane, position p = lane*EPL + e, ascending keys, dense from 0 ----
template <int EPL>
__device__ __forceinline__ void list_insert(uint32_t (&K)[EPL], uint32_t (&V)[EPL], uint32_t d, uint32_t nv, bool lane0) {
  uint32_t pk = __shfl_up_sync(FULL_MASK, K[EPL - 1], 1);
  const uint32_t pv = __shfl_up_sync(FULL_MASK, V[EPL - 1], 1);
  if (lane0) pk = 0u;  // nothing below position 0: "the entry below stays"
  // Slot e keeps its entry while key <= d (the new one lands after its equals); otherwise it takes the new
  // entry if the slot below keeps its own, else the entry of the slot below.  One predicated select per
  // register, highest slot first so that every source is still the old value.
#pragma unroll
  for (int e = EPL - 1; e >= 0; e--) {
    const uint32_t prk = e > 0 ? K[e - 1] : pk;
    const uint32_t prv = e > 0 ? V[e - 1] : pv;
    asm("{\n"
        ".reg .pred stay, here;\n"
        "setp.le.u32 stay, %0, %2;\n"
        "setp.le.u32 here, %3, %2;\n"
        "@!stay selp.u32 %0, %2, %3, here;\n"
        "@!stay selp.u32 %1, %4, %5, here;\n"
        "}\n"
        : "+r"(K[e]), "+r"(V[e])
        : "r"(d), "r"(prk), "r"(nv), "r"(prv));
  }
}

template <int EPL>
__device__ __forceinline__ uint32_t list_key_at(const uint32_t (&K)[EPL], int lane_of, int slot_of) {
  uint32_t sel = K[0];
#pragma unroll
  for (int e = 1; e < EPL; e++)
    if (slot_of == e) sel = K[e];
  return __shfl_sync(FULL_MASK, sel, lane_of);
}

}  // namespace
