// Division by a runtime-invariant 32-bit divisor with a precomputed magic multiplier (one IMAD.HI + shift instead of
// the ~20-instruction software divide; 64-bit divides cost > 100 instructions and dominated the index arithmetic of the
// issue-bound NHWC layer kernels).
#pragma once
#include <cstdint>

namespace psd {

struct FastDiv {
  uint32_t mul, shift, d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  // exact for all 32-bit n when computed with a 64-bit product (n < 2^31 here)
  FastDiv f;
  f.d = d;
  if (d == 1) { f.mul = 0; f.shift = 0; return f; }
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  f.shift = s;
  f.mul = static_cast<uint32_t>(((1ull << (32 + s)) + d - 1) / d - (1ull << 32));
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  if (f.d == 1) return n;
  const uint32_t hi = __umulhi(n, f.mul);
  return (hi + ((n - hi) >> 1)) >> (f.shift - 1);
}
__device__ __forceinline__ uint32_t fmod_(uint32_t n, uint32_t q, const FastDiv& f) { return n - q * f.d; }

}  // namespace psd
