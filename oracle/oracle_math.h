/*
 * oracle/oracle_math.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * "Platform" arithmetic of the reference shaders, frozen.
 *
 * The reference (apbodnar/FSPT) leaves sin/cos/atan/asin/pow/min/max/normalize
 * to whatever GLSL ES 3.00 implementation runs the shaders.  Nothing in the
 * reference pins them, so the oracle pins them itself, as spec "FSPT-DM1":
 *
 *   - every f32 operation is a single IEEE-754 binary32 round-to-nearest-even
 *     operation, no contraction (compile with -ffp-contract=off, no fast-math);
 *   - transcendental built-ins are evaluated in IEEE binary64 by the fixed
 *     operation sequences below (only + - * / sqrt rint on doubles) and rounded
 *     ONCE to binary32;
 *   - min/max are IEEE-754 minNum/maxNum (what GPU FMNMX does; GLSL leaves the
 *     NaN case undefined);
 *   - vector built-ins follow the formulas printed in the GLSL ES 3.00 spec
 *     section 8 (normalize = x/length(x), reflect, refract, mix, clamp, fract).
 *
 * The CUDA product implements the same spec independently in
 * fspt_b200/csrc/dm_math.cuh; tests/ compare the two bit for bit.
 */
#ifndef FSPT_ORACLE_MATH_H
#define FSPT_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace om {

/* ---- scalar helpers --------------------------------------------------- */
static inline float fminN(float a, float b) { /* IEEE minNum */
  if (a != a) return b;
  if (b != b) return a;
  return b < a ? b : a;
}
static inline float fmaxN(float a, float b) { /* IEEE maxNum */
  if (a != a) return b;
  if (b != b) return a;
  return a < b ? b : a;
}
static inline float clampf(float x, float lo, float hi) { return fminN(fmaxN(x, lo), hi); }
static inline float fractf(float x) { return x - floorf(x); }
static inline float mixf(float x, float y, float a) { return x * (1.0f - a) + y * a; }

/* ---- FSPT-DM1 transcendental kernels (binary64) ------------------------ */
static const double DM_TWO_OVER_PI = 0x1.45f306dc9c883p-1;
static const double DM_PIO2_A = 0x1.921fb54000000p+0;  /* top 27 bits of pi/2 */
static const double DM_PIO2_B = 0x1.10b4610000000p-30; /* next 27 bits        */
static const double DM_PIO2_C = 0x1.a62633145c06ep-58; /* remainder           */
static const double DM_PI = 0x1.921fb54442d18p+1;
static const double DM_PIO2 = 0x1.921fb54442d18p+0;
static const double DM_PIO4 = 0x1.921fb54442d18p-1;
static const double DM_LN2 = 0x1.62e42fefa39efp-1;
static const double DM_LOG2E = 0x1.71547652b82fep+0;

/* sin/cos on |r| <= pi/4 : fdlibm k_sin/k_cos minimax coefficients, Horner */
static inline double dm_ksin(double r) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = r * r;
  double v = z * r;
  double p = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return r + v * (S1 + z * p);
}
static inline double dm_kcos(double r) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = r * r;
  double p = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  return 1.0 - (0.5 * z - z * p);
}
/* range reduction: r = x - k*pi/2, k = rint(x*2/pi); accurate for |x| < 2^26 */
static inline double dm_reduce(double x, int64_t* kq) {
  double k = rint(x * DM_TWO_OVER_PI);
  double r = ((x - k * DM_PIO2_A) - k * DM_PIO2_B) - k * DM_PIO2_C;
  /* beyond the accurate domain keep the result bounded and deterministic */
  if (!(k > -9.0e15 && k < 9.0e15)) { *kq = 0; return 0.0; }
  *kq = (int64_t)k;
  return r;
}
static inline float dm_sin(float x) {
  int64_t k;
  double r = dm_reduce((double)x, &k);
  double s;
  switch ((int)(k & 3)) {
    case 0: s = dm_ksin(r); break;
    case 1: s = dm_kcos(r); break;
    case 2: s = -dm_ksin(r); break;
    default: s = -dm_kcos(r); break;
  }
  return (float)s;
}
static inline float dm_cos(float x) {
  int64_t k;
  double r = dm_reduce((double)x, &k);
  double s;
  switch ((int)(k & 3)) {
    case 0: s = dm_kcos(r); break;
    case 1: s = -dm_ksin(r); break;
    case 2: s = -dm_kcos(r); break;
    default: s = dm_ksin(r); break;
  }
  return (float)s;
}

/* atan on [0,1]: one reduction about tan(pi/8), then odd Taylor to a^21 */
static inline double dm_atan01(double a) {
  double base = 0.0;
  if (a > 0.41421356237309503) { /* tan(pi/8) */
    a = (a - 1.0) / (a + 1.0);
    base = DM_PIO4;
  }
  double z = a * a;
  double p = 1.0 / 21.0;
  p = -1.0 / 19.0 + z * p;
  p = 1.0 / 17.0 + z * p;
  p = -1.0 / 15.0 + z * p;
  p = 1.0 / 13.0 + z * p;
  p = -1.0 / 11.0 + z * p;
  p = 1.0 / 9.0 + z * p;
  p = -1.0 / 7.0 + z * p;
  p = 1.0 / 5.0 + z * p;
  p = -1.0 / 3.0 + z * p;
  p = 1.0 + z * p;
  return base + a * p;
}
static inline double dm_atan2d(double y, double x) {
  double ax = fabs(x), ay = fabs(y);
  double hi = ax > ay ? ax : ay;
  double lo = ax > ay ? ay : ax;
  if (!(hi > 0.0)) return 0.0; /* atan(0,0) and NaN inputs: defined as 0 */
  double r = dm_atan01(lo / hi);
  if (ay > ax) r = DM_PIO2 - r;
  if (x < 0.0) r = DM_PI - r;
  if (y < 0.0) r = -r;
  return r;
}
static inline float dm_atan2(float y, float x) { return (float)dm_atan2d((double)y, (double)x); }
static inline float dm_asin(float x) {
  double xd = (double)x;
  if (xd > 1.0) xd = 1.0;   /* |x| may exceed 1 by an ulp after normalize() */
  if (xd < -1.0) xd = -1.0;
  return (float)dm_atan2d(xd, sqrt((1.0 - xd) * (1.0 + xd)));
}

/* 2^x for double x; Taylor of e^t, t = f*ln2, |f| <= 0.5, to t^12 */
static inline double dm_exp2d(double x) {
  if (x != x) return x;
  if (x > 1000.0) x = 1000.0;
  if (x < -1100.0) x = -1100.0;
  double n = rint(x);
  double t = (x - n) * DM_LN2;
  double p = 1.0 / 479001600.0;
  p = 1.0 / 39916800.0 + t * p;
  p = 1.0 / 3628800.0 + t * p;
  p = 1.0 / 362880.0 + t * p;
  p = 1.0 / 40320.0 + t * p;
  p = 1.0 / 5040.0 + t * p;
  p = 1.0 / 720.0 + t * p;
  p = 1.0 / 120.0 + t * p;
  p = 1.0 / 24.0 + t * p;
  p = 1.0 / 6.0 + t * p;
  p = 0.5 + t * p;
  p = 1.0 + t * p;
  p = 1.0 + t * p;
  /* scale by 2^n in two exact steps (n in [-1100,1000]) */
  int ni = (int)n;
  int n1 = ni / 2, n2 = ni - n1;
  uint64_t b1 = (uint64_t)(int64_t)(n1 + 1023) << 52, b2 = (uint64_t)(int64_t)(n2 + 1023) << 52;
  double s1, s2;
  memcpy(&s1, &b1, 8);
  memcpy(&s2, &b2, 8);
  return (p * s1) * s2;
}
/* log2 for double x > 0 (normal): atanh series to s^17 */
static inline double dm_log2d(double x) {
  uint64_t b;
  memcpy(&b, &x, 8);
  int e = (int)((b >> 52) & 0x7ff) - 1023;
  b = (b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
  double m;
  memcpy(&m, &b, 8);
  if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  double p = 1.0 / 17.0;
  p = 1.0 / 15.0 + z * p;
  p = 1.0 / 13.0 + z * p;
  p = 1.0 / 11.0 + z * p;
  p = 1.0 / 9.0 + z * p;
  p = 1.0 / 7.0 + z * p;
  p = 1.0 / 5.0 + z * p;
  p = 1.0 / 3.0 + z * p;
  p = 1.0 + z * p;
  return (double)e + (2.0 * s * p) * DM_LOG2E;
}
static inline float dm_exp2(float x) { return (float)dm_exp2d((double)x); }
/* pow(x,y) = exp2(y*log2(x)) (GLSL ES 3.00 8.2); x <= 0 or NaN -> 0 (GLSL: undefined) */
static inline float dm_pow(float x, float y) {
  if (!(x > 0.0f)) return 0.0f;
  if (x > 3.0e38f) return x;
  return (float)dm_exp2d((double)y * dm_log2d((double)x));
}

/* ---- vec types ---------------------------------------------------------- */
struct v2 { float x, y; };
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };

static inline v3 mk3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 mul(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v3 mul(float s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline v3 div(v3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 x, v3 y) { /* GLSL ES 3.00 8.4 */
  return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalize(v3 a) { return div(a, length(a)); }
static inline v3 reflect(v3 I, v3 N) { return sub(I, mul(2.0f * dot(N, I), N)); }
static inline v3 refract(v3 I, v3 N, float eta) {
  float d = dot(N, I);
  float k = 1.0f - eta * eta * (1.0f - d * d);
  if (k < 0.0f) return mk3(0.0f, 0.0f, 0.0f);
  return sub(mul(eta, I), mul(eta * d + sqrtf(k), N));
}
static inline v3 mix3(v3 x, v3 y, float a) { return mk3(mixf(x.x, y.x, a), mixf(x.y, y.y, a), mixf(x.z, y.z, a)); }
static inline v3 clamp3(v3 a, float lo, float hi) { return mk3(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)); }

}  // namespace om
#endif
