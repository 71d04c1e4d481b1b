// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory): bfloat16 storage type with round-to-nearest-even.
#pragma once
#include <cstdint>
#include <cstring>
struct __nv_bfloat16 { uint16_t x; };
static inline __nv_bfloat16 __float2bfloat16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  __nv_bfloat16 r;
  if ((u & 0x7fffffffu) > 0x7f800000u) { r.x = 0x7fff; return r; }       // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  r.x = (uint16_t)(u >> 16);
  return r;
}
static inline float __bfloat162float(__nv_bfloat16 h) { uint32_t u = (uint32_t)h.x << 16; float f; memcpy(&f, &u, 4); return f; }
struct __nv_bfloat162 { __nv_bfloat16 x, y; };      // .x = low half-word
static inline __nv_bfloat162 __floats2bfloat162_rn(float lo, float hi) { return __nv_bfloat162{__float2bfloat16(lo), __float2bfloat16(hi)}; }
