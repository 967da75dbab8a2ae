// Exactness of fspt_b200/csrc/fastdiv.h (compiled and run by tests/test_fastdiv.py; no GPU).
#include <cstdio>
#include <cstdlib>
#include <initializer_list>

#include "../fspt_b200/csrc/fastdiv.h"

int main() {
  unsigned seed = 777;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed; };
  long long checked = 0;
  int n_mul = 0, n_shift = 0, n_plain = 0;
  // divisors the kernels use: samples per wave (1..64), tiles per row / row widths (1..2^14 and a few large ones)
  for (unsigned d = 1; d <= 20000; d += (d < 4200 ? 1 : 37)) {
    for (unsigned long long max_n : {(1ull << 26) - 1, (1ull << 22) - 1, (1ull << 27) - 1, 1000ull}) {
      const FastDiv f = make_fastdiv(d, max_n);
      (f.mode >= 0 ? n_shift : f.mode == -1 ? n_mul : n_plain)++;
      auto check = [&](unsigned long long n) {
        if (n > max_n) return true;
        ++checked;
        if (fast_div((unsigned)n, f) != (unsigned)(n / d)) {
          fprintf(stderr, "d=%u n=%llu mode=%d: %u != %llu\n", d, n, f.mode, fast_div((unsigned)n, f), n / d);
          return false;
        }
        return true;
      };
      // multiples of d and their neighbours (where a rounding error would show), the top of the range, random values
      for (int k = 0; k < 200; ++k) {
        const unsigned long long q = (k < 100) ? (unsigned long long)k : (max_n / d) - (unsigned long long)(k - 100);
        for (long long off = -1; off <= 1; ++off) {
          const long long n = (long long)(q * d) + off;
          if (n >= 0 && !check((unsigned long long)n)) return 1;
        }
      }
      for (int k = 0; k < 300; ++k) if (!check(rnd() % (max_n + 1))) return 1;
      if (!check(max_n) || !check(max_n - 1) || !check(0)) return 1;
    }
  }
  // exhaustive for the two divisors of the bench frame (64 samples per wave, 160 tiles per row of a 1280-pixel frame)
  for (unsigned d : {64u, 160u, 48u, 7u}) {
    const FastDiv f = make_fastdiv(d, (1ull << 26) - 1);
    for (unsigned n = 0; n < (1u << 26); n += 1) {
      if (fast_div(n, f) != n / d) { fprintf(stderr, "exhaustive d=%u n=%u\n", d, n); return 1; }
    }
    checked += 1ll << 26;
  }
  printf("ok %lld divisions (%d shift, %d multiply, %d plain divisors)\n", checked, n_shift, n_mul, n_plain);
  return 0;
}
