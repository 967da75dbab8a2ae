// dm_math.cuh -- device implementation of the "FSPT-DM1" arithmetic model.
//
// The reference shaders leave sin/cos/atan/asin/pow to the GLSL platform (tracer.fs:181,412,417,429-430,
// camera.fs:19,27-34, draw.fs:92).  To make the CUDA path reproducible against the CPU oracle, these built-ins
// are evaluated here in IEEE binary64 with explicitly rounded, never-contracted operations (__dmul_rn /
// __dadd_rn / __ddiv_rn / __dsqrt_rn) and rounded once to binary32.  B200 issues 64 FP64 ops/clk/SM, so the
// cost is a few percent of a shading pass and nothing in traversal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dm {

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }

#define DM_TWO_OVER_PI 0x1.45f306dc9c883p-1
#define DM_PIO2_A 0x1.921fb54000000p+0
#define DM_PIO2_B 0x1.10b4610000000p-30
#define DM_PIO2_C 0x1.a62633145c06ep-58
#define DM_PI 0x1.921fb54442d18p+1
#define DM_PIO2 0x1.921fb54442d18p+0
#define DM_PIO4 0x1.921fb54442d18p-1
#define DM_LN2 0x1.62e42fefa39efp-1
#define DM_LOG2E 0x1.71547652b82fep+0

// Horner step  c + z*p  with separately rounded multiply and add
__device__ __forceinline__ double hs(double c, double z, double p) { return add(c, mul(z, p)); }

__device__ __forceinline__ double ksin(double r) {
  const double z = mul(r, r);
  const double v = mul(z, r);
  double p = hs(-2.50507602534068634195e-08, z, 1.58969099521155010221e-10);
  p = hs(2.75573137070700676789e-06, z, p);
  p = hs(-1.98412698298579493134e-04, z, p);
  p = hs(8.33333333332248946124e-03, z, p);
  return add(r, mul(v, hs(-1.66666666666666324348e-01, z, p)));
}
__device__ __forceinline__ double kcos(double r) {
  const double z = mul(r, r);
  double p = hs(2.08757232129817482790e-09, z, -1.13596475577881948265e-11);
  p = hs(-2.75573143513906633035e-07, z, p);
  p = hs(2.48015872894767294178e-05, z, p);
  p = hs(-1.38888888888741095749e-03, z, p);
  p = hs(4.16666666666666019037e-02, z, p);
  p = mul(z, p);
  return sub(1.0, sub(mul(0.5, z), mul(z, p)));
}
// r = x - k*pi/2 with a 27+27+53-bit split of pi/2 (k*A and k*B exact for |k| < 2^26)
__device__ __forceinline__ double reduce(double x, int& q) {
  const double k = rint(mul(x, DM_TWO_OVER_PI));
  const double r = sub(sub(sub(x, mul(k, DM_PIO2_A)), mul(k, DM_PIO2_B)), mul(k, DM_PIO2_C));
  if (!(k > -9.0e15 && k < 9.0e15)) { q = 0; return 0.0; }
  q = (int)(((long long)k) & 3);
  return r;
}
// both kernels are always evaluated and the quadrant only selects: no divergence inside a warp
__device__ __forceinline__ float sinf_(float x) {
  int q;
  const double r = reduce((double)x, q);
  const double a = ksin(r), b = kcos(r);
  double s = (q & 1) ? b : a;
  if (q & 2) s = -s;
  return (float)s;
}
__device__ __forceinline__ float cosf_(float x) {
  int q;
  const double r = reduce((double)x, q);
  const double a = ksin(r), b = kcos(r);
  double s = (q & 1) ? a : b;
  if (q == 1 || q == 2) s = -s;
  return (float)s;
}
// both at once (sampleMicrofacet, sampleLambert, sampleEnv, camera use the pair on one angle)
__device__ __forceinline__ void sincosf_(float x, float& sn, float& cs) {
  int q;
  const double r = reduce((double)x, q);
  const double a = ksin(r), b = kcos(r);
  double s = (q & 1) ? b : a;
  double c = (q & 1) ? a : b;
  if (q & 2) s = -s;
  if (q == 1 || q == 2) c = -c;
  sn = (float)s;
  cs = (float)c;
}

__device__ __forceinline__ double atan01(double a) {
  double base = 0.0;
  if (a > 0.41421356237309503) {
    a = div(sub(a, 1.0), add(a, 1.0));
    base = DM_PIO4;
  }
  const double z = mul(a, a);
  double p = 1.0 / 21.0;
  p = hs(-1.0 / 19.0, z, p);
  p = hs(1.0 / 17.0, z, p);
  p = hs(-1.0 / 15.0, z, p);
  p = hs(1.0 / 13.0, z, p);
  p = hs(-1.0 / 11.0, z, p);
  p = hs(1.0 / 9.0, z, p);
  p = hs(-1.0 / 7.0, z, p);
  p = hs(1.0 / 5.0, z, p);
  p = hs(-1.0 / 3.0, z, p);
  p = hs(1.0, z, p);
  return add(base, mul(a, p));
}
__device__ __forceinline__ double atan2d(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const double hi = ax > ay ? ax : ay;
  const double lo = ax > ay ? ay : ax;
  if (!(hi > 0.0)) return 0.0;
  double r = atan01(div(lo, hi));
  if (ay > ax) r = sub(DM_PIO2, r);
  if (x < 0.0) r = sub(DM_PI, r);
  if (y < 0.0) r = -r;
  return r;
}
__device__ __forceinline__ float atan2f_(float y, float x) { return (float)atan2d((double)y, (double)x); }
__device__ __forceinline__ float asinf_(float x) {
  double xd = (double)x;
  if (xd > 1.0) xd = 1.0;
  if (xd < -1.0) xd = -1.0;
  return (float)atan2d(xd, __dsqrt_rn(mul(sub(1.0, xd), add(1.0, xd))));
}

__device__ __forceinline__ double exp2d(double x) {
  if (x != x) return x;
  if (x > 1000.0) x = 1000.0;
  if (x < -1100.0) x = -1100.0;
  const double n = rint(x);
  const double t = mul(sub(x, n), DM_LN2);
  double p = 1.0 / 479001600.0;
  p = hs(1.0 / 39916800.0, t, p);
  p = hs(1.0 / 3628800.0, t, p);
  p = hs(1.0 / 362880.0, t, p);
  p = hs(1.0 / 40320.0, t, p);
  p = hs(1.0 / 5040.0, t, p);
  p = hs(1.0 / 720.0, t, p);
  p = hs(1.0 / 120.0, t, p);
  p = hs(1.0 / 24.0, t, p);
  p = hs(1.0 / 6.0, t, p);
  p = hs(0.5, t, p);
  p = hs(1.0, t, p);
  p = hs(1.0, t, p);
  const int ni = (int)n;
  const int n1 = ni / 2, n2 = ni - n1;
  const double s1 = __longlong_as_double((long long)(n1 + 1023) << 52);
  const double s2 = __longlong_as_double((long long)(n2 + 1023) << 52);
  return mul(mul(p, s1), s2);
}
__device__ __forceinline__ double log2d(double x) {
  long long b = __double_as_longlong(x);
  int e = (int)((b >> 52) & 0x7ff) - 1023;
  b = (b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL;
  double m = __longlong_as_double(b);
  if (m > 1.4142135623730951) { m = mul(m, 0.5); e += 1; }
  const double s = div(sub(m, 1.0), add(m, 1.0));
  const double z = mul(s, s);
  double p = 1.0 / 17.0;
  p = hs(1.0 / 15.0, z, p);
  p = hs(1.0 / 13.0, z, p);
  p = hs(1.0 / 11.0, z, p);
  p = hs(1.0 / 9.0, z, p);
  p = hs(1.0 / 7.0, z, p);
  p = hs(1.0 / 5.0, z, p);
  p = hs(1.0 / 3.0, z, p);
  p = hs(1.0, z, p);
  return add((double)e, mul(mul(mul(2.0, s), p), DM_LOG2E));
}
__device__ __forceinline__ float exp2f_(float x) { return (float)exp2d((double)x); }
__device__ __forceinline__ float powf_(float x, float y) {
  if (!(x > 0.0f)) return 0.0f;
  if (x > 3.0e38f) return x;
  return (float)exp2d(mul((double)y, log2d((double)x)));
}

}  // namespace dm
