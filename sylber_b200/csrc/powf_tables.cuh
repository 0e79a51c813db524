// fp64 replay of glibc's powf(x, 0.5f) for the on-device segmentation (see segment.cuh).
//
// Why: the reference computes cosine similarities with `np.float32_scalar ** .5`
// (sylber/utils/segment_utils.py:69 reached with 1-D arguments from :96 and :114), which NumPy evaluates with
// libm powf, not sqrtf.  glibc's powf is not correctly rounded: powf(x, .5f) != sqrtf(x) for 1 355 471 of the
// 2 130 706 432 positive normal floats (measured exhaustively).  Matching the reference's decisions bit for bit
// therefore needs the same algorithm: glibc 2.39 sysdeps/ieee754/flt-32/e_powf.c as compiled for the FMA ifunc
// variant (__powf_fma, chosen on any x86-64 with FMA+AVX2), i.e. log2 via a 16-entry table + degree-4
// polynomial, times y, then exp2 via a 32-entry table + cubic, all in double precision with the exact
// fused-multiply-add pattern of the shipped binary.  Constants below are that binary's tables
// (tools/extract_powf_tables.py re-reads them from libm.so.6).  The same sequence written in C with fma()
// equals powf(x, .5f) for every positive normal float (tests/test_powf_emulation.py).
#pragma once

#include <stdint.h>

namespace syl {

__device__ const double kPowfLogTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2,
    0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2,
    0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2,
    0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2,
    0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2,
    0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3,
    0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4,
    0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5,
    0x1.0000000000000p+0, 0x0.0p+0,
    0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4,
    0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3,
    0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3,
    0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2,
    0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2,
    0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2,
};
__device__ const double kPowfLogPoly[5] = {0x1.27616c9496e0bp-2, -0x1.71969a075c67ap-2, 0x1.ec70a6ca7baddp-2, -0x1.7154748bef6c8p-1, 0x1.71547652ab82bp+0};
__device__ const unsigned long long kExp2fTab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
};
constexpr double kExp2fShiftScaled = 0x1.8000000000000p+47;
__device__ const double kExp2fPolyScaled[3] = {0x1.c6af84b912394p-5, 0x1.ebfce50fac4f3p-3, 0x1.62e42ff0c52d6p-1};

// powf(x, 0.5f) for positive, normal, finite x (callers pass a sum of squares plus 1e-8)
__device__ __forceinline__ float powf_half(float x) {
  if (!(x < __int_as_float(0x7f800000))) return x;   // +inf -> +inf, NaN -> NaN, as powf
  const uint32_t ix = __float_as_uint(x);
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const uint32_t top = tmp & 0xff800000u;
  const int k = (int32_t)top >> 23;
  const double z = (double)__uint_as_float(ix - top);
  const double r = __fma_rn(z, kPowfLogTab[2 * i], -1.0);
  const double y0 = __dadd_rn(kPowfLogTab[2 * i + 1], (double)k);
  double y = __fma_rn(kPowfLogPoly[0], r, kPowfLogPoly[1]);
  const double p = __fma_rn(kPowfLogPoly[2], r, kPowfLogPoly[3]);
  const double r2 = __dmul_rn(r, r);
  double q = __fma_rn(r, kPowfLogPoly[4], y0);
  const double r4 = __dmul_rn(r2, r2);
  q = __fma_rn(r2, p, q);
  y = __fma_rn(y, r4, q);
  const double ylogx = __dmul_rn(0.5, y);
  double kd = __dadd_rn(ylogx, kExp2fShiftScaled);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, kExp2fShiftScaled);
  const double rr = __dsub_rn(ylogx, kd);
  const unsigned long long t = kExp2fTab[ki & 31] + (ki << 47);
  const double s = __longlong_as_double((long long)t);
  const double zz = __fma_rn(rr, kExp2fPolyScaled[0], kExp2fPolyScaled[1]);
  const double rr2 = __dmul_rn(rr, rr);
  double yy = __fma_rn(rr, kExp2fPolyScaled[2], 1.0);
  yy = __fma_rn(zz, rr2, yy);
  yy = __dmul_rn(yy, s);
  return __double2float_rn(yy);
}

}  // namespace syl
