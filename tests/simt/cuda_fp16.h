// TEST INFRASTRUCTURE: stand-in for <cuda_fp16.h> in the SIMT host emulation build -- IEEE binary16 through the compiler's _Float16.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
struct __half_raw {
  unsigned short x;
};
struct __half {
  uint16_t bits;
  __half() = default;
  explicit __half(uint16_t b) : bits(b) {}
  __half(const __half_raw &r) : bits(r.x) {}
};
inline __half __ushort_as_half(unsigned short u) { return __half((uint16_t)u); }
inline unsigned short __half_as_ushort(__half h) { return h.bits; }
inline float __half2float(__half h) {
  _Float16 f;
  std::memcpy(&f, &h.bits, 2);
  return (float)f;
}
inline __half __float2half_rn(float x) {
  const _Float16 f = (_Float16)x;
  __half h;
  std::memcpy(&h.bits, &f, 2);
  return h;
}
inline __half __float2half_rd(float x) {  // round toward -infinity
  __half h = __float2half_rn(x);
  if (__half2float(h) > x) {  // step to the next representable value below
    if (h.bits == 0x0000) h.bits = 0x8001;              // +0 -> smallest negative subnormal
    else if (h.bits & 0x8000) h.bits += 1;              // negative: larger magnitude
    else h.bits -= 1;                                   // positive: smaller magnitude
  }
  return h;
}
