// libm_sincosf.cuh — sinf/cosf with the bits of the host C library the reference runs on.
//
// computeOrbDescriptor (reference src/ORBextractor.cpp:116) takes `(float)cos(angle)` and
// `(float)sin(angle)` of a float angle: the std::cos(float)/std::sin(float) overloads, i.e. glibc's
// cosf/sinf.  Those are NOT correctly rounded (0.3-0.6 % of the arguments in [0, 2*pi] differ from
// the rounded double result, measured), and the rotated sampling pattern is rounded to pixels from
// them, so a descriptor that has to be bit-exact needs the same function, not a better one.
//
// This restates the published algorithm of glibc >= 2.28 sysdeps/ieee754/flt-32/{s_sinf,s_cosf,
// sincosf.h} (the ARM optimized-routines sinf/cosf): argument widened to double, fast reduction by
// pi/2 for |x| < 120, degree-7 / degree-8 minimax polynomials, one rounding to float.  The operation
// order and the places where a multiply-add is fused follow the variant glibc selects on every
// x86-64 CPU with FMA (__sinf_fma / __cosf_fma).  Pinned exhaustively on the host against libm for
// every float in [0, 7] (tests/test_orb_cpu.py) and on the device against the same (tests/gpu).
//
// Domain: |x| < 120 (the descriptor stage only passes angles in [0, 2*pi]); outside it the device
// function traps -- it is not a general sinf.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDA_ARCH__)
#define LORB_SC_HD __device__ __forceinline__
#define LORB_SC_MUL(a, b) __dmul_rn((a), (b))
#define LORB_SC_FMA(a, b, c) __fma_rn((a), (b), (c))
#define LORB_SC_D2I(a) __double2int_rz(a)
#define LORB_SC_D2F(a) __double2float_rn(a)
#define LORB_SC_BITS(f) __float_as_uint(f)
#else
#define LORB_SC_HD static inline
#define LORB_SC_MUL(a, b) ((a) * (b))
#define LORB_SC_FMA(a, b, c) fma((a), (b), (c))
#define LORB_SC_D2I(a) ((int32_t)(a))
#define LORB_SC_D2F(a) ((float)(a))
static inline uint32_t lorb_sc_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
#define LORB_SC_BITS(f) lorb_sc_bits(f)
#endif

namespace lorb_libm {

// __sincosf_table[0]; table[1] is the same with the cosine coefficients negated.
#define LORB_SC_HPI_INV 0x1.45F306DC9C883p+23 /* 2/pi * 2^24 */
#define LORB_SC_HPI 0x1.921FB54442D18p0
#define LORB_SC_C0 0x1p0
#define LORB_SC_C1 -0x1.ffffffd0c621cp-2
#define LORB_SC_C2 0x1.55553e1068f19p-5
#define LORB_SC_C3 -0x1.6c087e89a359dp-10
#define LORB_SC_C4 0x1.99343027bf8c3p-16
#define LORB_SC_S1 -0x1.555545995a603p-3
#define LORB_SC_S2 0x1.1107605230bc4p-7
#define LORB_SC_S3 -0x1.994eb3774cf24p-13

// sinf_poly of sincosf.h: sine polynomial for even n, cosine polynomial (sign csgn) for odd n.
LORB_SC_HD float poly(double x, double x2, double csgn, int n) {
  if ((n & 1) == 0) {
    const double x3 = LORB_SC_MUL(x, x2);
    const double s1 = LORB_SC_FMA(LORB_SC_S3, x2, LORB_SC_S2);
    const double x5 = LORB_SC_MUL(x2, x3);
    const double s = LORB_SC_FMA(x3, LORB_SC_S1, x);
    return LORB_SC_D2F(LORB_SC_FMA(s1, x5, s));
  }
  const double x4 = LORB_SC_MUL(x2, x2);
  const double c1 = LORB_SC_FMA(csgn * LORB_SC_C1, x2, csgn * LORB_SC_C0);
  const double c2 = LORB_SC_FMA(csgn * LORB_SC_C4, x2, csgn * LORB_SC_C3);
  const double x6 = LORB_SC_MUL(x2, x4);
  const double c = LORB_SC_FMA(x4, csgn * LORB_SC_C2, c1);
  return LORB_SC_D2F(LORB_SC_FMA(c2, x6, c));
}

// is_cos = 0: sinf(y); 1: cosf(y).  |y| < 120.
LORB_SC_HD float sincosf_one(float y, int is_cos) {
  const uint32_t top = (LORB_SC_BITS(y) >> 20) & 0x7ff;
  double x = (double)y;
  if (top < 0x3f4) {  // |y| < pi/4
    const double x2 = LORB_SC_MUL(x, x);
    if (top < 0x398) return is_cos ? 1.0f : y;  // |y| < 2^-12
    return poly(x, x2, 1.0, is_cos);
  }
#if defined(__CUDA_ARCH__)
  if (top >= 0x42f) __trap();
#endif
  const double r = LORB_SC_MUL(x, LORB_SC_HPI_INV);
  const int n = (LORB_SC_D2I(r) + 0x800000) >> 24;
  x = LORB_SC_FMA(-(double)n, LORB_SC_HPI, x);
  const double sgn = ((n + 1) & 2) ? -1.0 : 1.0;  // sign[n & 3] = {1, -1, -1, 1}
  const double csgn = (n & 2) ? -1.0 : 1.0;
  return poly(LORB_SC_MUL(x, sgn), LORB_SC_MUL(x, x), csgn, n ^ is_cos);
}

LORB_SC_HD float sinf_libm(float y) { return sincosf_one(y, 0); }
LORB_SC_HD float cosf_libm(float y) { return sincosf_one(y, 1); }

}  // namespace lorb_libm
