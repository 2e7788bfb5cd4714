// Short fp64 transcendental kernels for the step path (sm_100a).
//
// CUDA's log() / sincos() cost 85-100 and ~77 SASS instructions per call here
// (profiles/r1_notes.md): range handling for denormals, huge arguments, Payne-Hanek
// reduction, and hi/lo compensation for a <1 ulp guarantee.  The step path calls
// log 29x and sincos 12x per env-step on arguments whose range is known, and the
// values feed quantities with a 1e-9 relative bar (rewards, free-running noise), so
// ~1-2 ulp kernels without the special cases are used instead.  Explicit fma() is
// honoured even though the translation unit is compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>

namespace mdg {

// 1/d for normal finite d: MUFU.RCP64H seed (~20 bits) + two Newton steps
__device__ __forceinline__ double fast_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// sqrt(x), x positive, finite and normal; ~1 ulp: MUFU.RSQ64H seed (~22 bits) + two coupled Newton steps on
// (g ~ sqrt x, h ~ 1/(2 sqrt x)).  8 instructions instead of the ~19 of the IEEE sqrt sequence.
__device__ __forceinline__ double fast_sqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  return fma(g, r, g);
}

// log(x), x positive, finite and normal; ~1-2 ulp.  x = 2^e * m, m in [sqrt(1/2), sqrt(2)),
// log m = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716: odd series through s^19 (next term 2.3e-17).
__device__ __forceinline__ double fast_log_pos(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000FFFFF) | 0x3FF00000;
  if (hi >= 0x3FF6A09F) {  // m >= ~sqrt(2): halve it
    hi -= 0x00100000;
    e += 1;
  }
  const double m = __hiloint2double(hi, lo);
  const double f = m - 1.0;
  const double s = f * fast_rcp(2.0 + f);
  const double z = s * s;
#ifdef MDG_ESTRIN
  // Estrin's scheme: depth 4 instead of 8 dependent fma
  const double z2 = z * z, z4 = z2 * z2;
  const double a0 = fma(2.0 / 5.0, z, 2.0 / 3.0), a1 = fma(2.0 / 9.0, z, 2.0 / 7.0);
  const double a2 = fma(2.0 / 13.0, z, 2.0 / 11.0), a3 = fma(2.0 / 17.0, z, 2.0 / 15.0);
  const double b0 = fma(a1, z2, a0), b1 = fma(a3, z2, a2);
  double p = fma(b1, z4, b0);
  p = fma((2.0 / 19.0) * z4, z4, p);
#else
  double p = 2.0 / 19.0;
  p = fma(p, z, 2.0 / 17.0);
  p = fma(p, z, 2.0 / 15.0);
  p = fma(p, z, 2.0 / 13.0);
  p = fma(p, z, 2.0 / 11.0);
  p = fma(p, z, 2.0 / 9.0);
  p = fma(p, z, 2.0 / 7.0);
  p = fma(p, z, 2.0 / 5.0);
  p = fma(p, z, 2.0 / 3.0);
#endif
  const double logm = fma(s * z, p, 2.0 * s);
  const double de = (double)e;
  // ln2 split so that e*ln2_hi is exact for |e| < 2^11
  return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, logm));
}

// log(x) for any x: the short kernel on its domain, CUDA's log() (one out-of-line copy) elsewhere
// (0, negative, inf, NaN, denormal)
static __device__ __noinline__ double slow_log(double x) { return log(x); }
__device__ __forceinline__ double fast_log(double x) {
  if (x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308) return fast_log_pos(x);
  return slow_log(x);
}

// sin(2*pi*u), cos(2*pi*u) for u in [0,1): exact quadrant reduction (4u - rint(4u) is exact),
// Taylor kernels on |x| <= pi/4 through x^15 / x^16 (truncation 5e-17 / 2e-18).
__device__ __forceinline__ void fast_sincos_2pi(double u, double& sn, double& cs) {
  const double t = 4.0 * u;
  const double q = rint(t);
  const double x = (t - q) * 1.57079632679489661923;
  const double x2 = x * x;
  double s = -7.6471637318198164759e-13;        // -1/15!
  s = fma(s, x2, 1.6059043836821614599e-10);    //  1/13!
  s = fma(s, x2, -2.5052108385441718775e-08);   // -1/11!
  s = fma(s, x2, 2.7557319223985890653e-06);    //  1/9!
  s = fma(s, x2, -1.9841269841269841270e-04);   // -1/7!
  s = fma(s, x2, 8.3333333333333333333e-03);    //  1/5!
  s = fma(s, x2, -1.6666666666666666667e-01);   // -1/3!
  const double sx = fma(x * x2, s, x);
  double c = 4.7794773323873852974e-14;         //  1/16!
  c = fma(c, x2, -1.1470745597729724714e-11);   // -1/14!
  c = fma(c, x2, 2.0876756987868098979e-09);    //  1/12!
  c = fma(c, x2, -2.7557319223985890653e-07);   // -1/10!
  c = fma(c, x2, 2.4801587301587301587e-05);    //  1/8!
  c = fma(c, x2, -1.3888888888888888889e-03);   // -1/6!
  c = fma(c, x2, 4.1666666666666666667e-02);    //  1/4!
  c = fma(c, x2, -0.5);                          // -1/2!
  const double cx = fma(c, x2, 1.0);
  const int k = (int)q & 3;
  const double a = (k & 1) ? cx : sx;   // sin candidate
  const double b = (k & 1) ? sx : cx;   // cos candidate
  sn = (k & 2) ? -a : a;
  cs = ((k + 1) & 2) ? -b : b;
}

}  // namespace mdg
