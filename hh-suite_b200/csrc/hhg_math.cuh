// hhg_math.cuh -- the reference's approximate log2 / pow2 (src/util-inl.h:83-214) as device functions, bit for bit.
// No other dependencies: shared by the Viterbi / prefilter kernels, the HHM loader and the alignment -> HMM kernels
// (and by the CPU emulation of the latter under tests/emul).
#pragma once
#include <cfloat>
#include <cstdint>

namespace hhg {

// fast_log2 (src/util-inl.h:108-128): table lookup + linear interpolation, x > 0 else -100000
__device__ __forceinline__ float fast_log2_dev(float x, const float* lg2, const float* diff) {
  if (!(x > 0.0f)) return -100000.0f;
  const uint32_t u = __float_as_uint(x);
  const int a = (int)((u & 0x7F800000u) >> 23) - 0x7f;
  const int b = (int)((u & 0x007FE000u) >> 13);
  const int c = (int)(u & 0x00001FFFu);
  return __fadd_rn(__fadd_rn((float)a, lg2[b]), __fmul_rn(diff[b], (float)c));
}

// flog2, src/util-inl.h:83-93 (the polynomial is evaluated in double, as the C++ expression promotes)
__device__ __forceinline__ float flog2_dev(float x) {
  if (x <= 0.f) return -128.f;
  uint32_t u = __float_as_uint(x);
  const float e = (float)((int)((u & 0x7F800000u) >> 23) - 0x7f);
  x = __uint_as_float((u & 0x007FFFFFu) | 0x3f800000u);
  x = __double2float_rn(__dsub_rn((double)x, 1.0));
  const double xd = (double)x;
  double y = __dadd_rn(-0.1903190, __dmul_rn(xd, 0.0440047));
  y = __dadd_rn(0.4123442, __dmul_rn(xd, y));
  y = __dadd_rn(-0.7077702, __dmul_rn(xd, y));
  y = __dadd_rn(1.441740, __dmul_rn(xd, y));
  x = __double2float_rn(__dmul_rn(xd, y));
  return __fadd_rn(x, e);
}

// fpow2, src/util-inl.h:190-214
__device__ __forceinline__ float fpow2_dev(float x) {
  if (x >= 128.0f) return FLT_MAX;
  if (x <= -125.0f) return 0.0f;
  const float tx = __fadd_rn(__fsub_rn(x, 0.5f), 12582912.0f);
  const int lx = __float_as_int(tx) - 0x4b400000;
  const float dx = __fsub_rn(x, (float)lx);
  float y = __fadd_rn(0.0520749f, __fmul_rn(dx, 0.0134929f));
  y = __fadd_rn(0.241404f, __fmul_rn(dx, y));
  y = __fadd_rn(0.693019f, __fmul_rn(dx, y));
  y = __fadd_rn(1.0f, __fmul_rn(dx, y));
  return __int_as_float(__float_as_int(y) + (lx << 23));
}

}  // namespace hhg
