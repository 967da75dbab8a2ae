// fastdiv.h -- n / d for a divisor that is fixed per launch (samples per wave, tiles per row), without the ~22-instruction
// integer-division sequence the kernels used to run two or three times per ray and per shaded vertex.
//   d a power of two          -> shift
//   otherwise                 -> floor(n * ceil(2^39 / d) / 2^39), exact while n * d <= 2^39 and the product fits 64 bits
//   else (huge frames)        -> plain division
// The host decides (make_fastdiv, given the largest dividend of the launch); tests/test_fastdiv.py checks exactness on the CPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FSPT_HD __host__ __device__ __forceinline__
#else
#define FSPT_HD inline
#endif

struct FastDiv {
  unsigned d;
  int mode;                // >= 0: d == 1 << mode;  -1: multiply-high;  -2: plain division
  unsigned long long mul;  // ceil(2^39 / d)
};

inline FastDiv make_fastdiv(unsigned d, unsigned long long max_n) {
  FastDiv f;
  f.d = d ? d : 1u;
  f.mul = 0;
  f.mode = -2;
  if ((f.d & (f.d - 1)) == 0) {
    int s = 0;
    while ((1u << s) < f.d) ++s;
    f.mode = s;
    return f;
  }
  const unsigned long long K = 1ull << 39;
  f.mul = (K + f.d - 1) / f.d;
  // exact for n <= max_n when max_n * d <= 2^39; the 64-bit product needs max_n * mul < 2^64
  if (max_n <= K / f.d && (max_n == 0 || f.mul <= ~0ull / max_n)) f.mode = -1;
  return f;
}

FSPT_HD unsigned fast_div(unsigned n, const FastDiv& f) {
  if (f.mode >= 0) return n >> f.mode;
  if (f.mode == -1) return (unsigned)(((unsigned long long)n * f.mul) >> 39);
  return n / f.d;
}
