// dm_math.cuh -- device implementation of the "FSPT-DM2" arithmetic model.
//
// The reference shaders leave sin/cos/atan/asin/pow to the GLSL platform (tracer.fs:181,412,417,429-430,
// camera.fs:19,27-34, draw.fs:92).  To make the CUDA path reproducible against the CPU oracle these built-ins are
// evaluated by fixed binary32 operation sequences: explicitly rounded add/mul/div/sqrt (never contracted) and
// explicit single-rounded FMAs (__fmaf_rn), which x86 FMA3 reproduces bit for bit on the oracle side.
// (DM1, the first version, evaluated them in binary64: 46 % of the shading kernel's instructions.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dm {

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }

#define DM_TWO_OVER_PI 0x1.45f306p-1f
#define DM_PIO2_1 0x1.921fb6p+0f
#define DM_PIO2_2 -0x1.777a5cp-25f
#define DM_PIO2_3 -0x1.ee59dap-50f
#define DM_PI 0x1.921fb6p+1f
#define DM_PIO2 0x1.921fb6p+0f
#define DM_PIO4 0x1.921fb6p-1f
#define DM_LN2 0x1.62e430p-1f
#define DM_LOG2E 0x1.715476p+0f

// r = x - k*pi/2: three-term Cody-Waite with exact products inside the FMAs, then one more fold because the f32
// product x*2/pi can miss the nearest integer by one for |x| > ~1e5 (the sin-hash RNG reaches 1e7)
__device__ __forceinline__ float reduce(float x, int& q) {
  if (!(fabsf(x) <= 1.0e9f)) { q = 0; return 0.0f; }
  const float k = rintf(mul_(x, DM_TWO_OVER_PI));
  float r = fma_(-k, DM_PIO2_1, x);
  r = fma_(-k, DM_PIO2_2, r);
  r = fma_(-k, DM_PIO2_3, r);
  const float k2 = rintf(mul_(r, DM_TWO_OVER_PI));
  r = fma_(-k2, DM_PIO2_1, r);
  r = fma_(-k2, DM_PIO2_2, r);
  q = ((int)k + (int)k2) & 3;
  return r;
}
__device__ __forceinline__ float ksin(float r) {
  const float z = mul_(r, r);
  float p = fma_(2.7234684694121825e-06f, z, -0.00019839966262225062f);
  p = fma_(p, z, 0.008333331905305386f);
  p = fma_(p, z, -0.1666666716337204f);
  return fma_(mul_(r, z), p, r);
}
__device__ __forceinline__ float kcos(float r) {
  const float z = mul_(r, r);
  float p = fma_(2.453538400004618e-05f, z, -0.001388824312016368f);
  p = fma_(p, z, 0.0416666641831398f);
  return fma_(mul_(z, z), p, fma_(-0.5f, z, 1.0f));
}
// both kernels are always evaluated and the quadrant only selects: no divergence inside a warp
__device__ __forceinline__ float sinf_(float x) {
  int q;
  const float r = reduce(x, q);
  const float a = ksin(r), b = kcos(r);
  const float s = (q & 1) ? b : a;
  return (q & 2) ? -s : s;
}
__device__ __forceinline__ float cosf_(float x) {
  int q;
  const float r = reduce(x, q);
  const float a = ksin(r), b = kcos(r);
  const float s = (q & 1) ? a : b;
  return (q == 1 || q == 2) ? -s : s;
}
// both at once (sampleMicrofacet, sampleLambert, sampleEnv, camera use the pair on one angle)
__device__ __forceinline__ void sincosf_(float x, float& sn, float& cs) {
  int q;
  const float r = reduce(x, q);
  const float a = ksin(r), b = kcos(r);
  const float s = (q & 1) ? b : a;
  const float c = (q & 1) ? a : b;
  sn = (q & 2) ? -s : s;
  cs = (q == 1 || q == 2) ? -c : c;
}

__device__ __forceinline__ float atan01(float a) {
  float base = 0.0f;
  if (a > 0.4142135679721832f) {
    a = div_(sub_(a, 1.0f), add_(a, 1.0f));
    base = DM_PIO4;
  }
  const float z = mul_(a, a);
  float p = fma_(-0.06418270617723465f, z, 0.10733865201473236f);
  p = fma_(p, z, -0.14263083040714264f);
  p = fma_(p, z, 0.19999517500400543f);
  p = fma_(p, z, -0.3333333134651184f);
  return add_(base, fma_(mul_(a, z), p, a));
}
__device__ __forceinline__ float atan2f_(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float hi = ax > ay ? ax : ay;
  const float lo = ax > ay ? ay : ax;
  if (!(hi > 0.0f)) return 0.0f;
  float r = atan01(div_(lo, hi));
  if (ay > ax) r = sub_(DM_PIO2, r);
  if (x < 0.0f) r = sub_(DM_PI, r);
  if (y < 0.0f) r = -r;
  return r;
}
__device__ __forceinline__ float asinf_(float x) {
  if (x > 1.0f) x = 1.0f;
  if (x < -1.0f) x = -1.0f;
  return atan2f_(x, __fsqrt_rn(mul_(sub_(1.0f, x), add_(1.0f, x))));
}

__device__ __forceinline__ float exp2f_(float x) {
  if (x != x) return x;
  if (x > 130.0f) x = 130.0f;
  if (x < -160.0f) x = -160.0f;
  const float n = rintf(x);
  const float f = sub_(x, n);
  float p = fma_(0.0001545316627016291f, f, 0.00133813067805022f);
  p = fma_(p, f, 0.009618083015084267f);
  p = fma_(p, f, 0.055503811687231064f);
  p = fma_(p, f, 0.24022650718688965f);
  const float r = fma_(mul_(f, f), p, fma_(f, DM_LN2, 1.0f));
  const int ni = (int)n;
  const int n1 = ni / 2, n2 = ni - n1;
  const float s1 = __int_as_float((n1 + 127) << 23), s2 = __int_as_float((n2 + 127) << 23);
  return mul_(mul_(r, s1), s2);
}
__device__ __forceinline__ float log2f_(float x) {
  int e = 0;
  if (x < 1.17549435e-38f) { x = mul_(x, 16777216.0f); e = -24; }
  unsigned b = __float_as_uint(x);
  e += (int)((b >> 23) & 0xffu) - 127;
  b = (b & 0x007fffffu) | 0x3f800000u;
  float m = __uint_as_float(b);
  if (m > 1.4142135381698608f) { m = mul_(m, 0.5f); e += 1; }
  const float s = div_(sub_(m, 1.0f), add_(m, 1.0f));
  const float z = mul_(s, s);
  float p = fma_(0.233596533536911f, z, 0.2855019271373749f);
  p = fma_(p, z, 0.4000011682510376f);
  p = fma_(p, z, 0.6666666865348816f);
  const float lnm = fma_(mul_(s, z), p, mul_(2.0f, s));
  return fma_(lnm, DM_LOG2E, (float)e);
}
__device__ __forceinline__ float powf_(float x, float y) {
  if (!(x > 0.0f)) return 0.0f;
  if (x > 3.0e38f) return x;
  return exp2f_(mul_(y, log2f_(x)));
}

}  // namespace dm
