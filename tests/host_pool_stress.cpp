// Stress test of fspt_b200/csrc/host_pool.h (compiled and run by tests/test_host_pool.py; no GPU).
#include <cstdio>
#include <cstdlib>
#include <numeric>

#include "../fspt_b200/csrc/host_pool.h"

int main() {
  std::atomic<int> inits(0);
  HostPool pool([&]() { inits.fetch_add(1); });
  unsigned seed = 12345;
  auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
  long long regions = 0;
  // regions of every shape from one thread: 0 / 1 / few / many items, 1 .. 24 workers, tiny and uneven items
  for (int rep = 0; rep < 4000; ++rep) {
    const int n = (int)(rnd() % 5 == 0 ? rnd() % 3 : rnd() % 300);
    const int w = 1 + (int)(rnd() % 24);
    std::vector<std::atomic<int>> hits((size_t)std::max(n, 1));
    for (auto& h : hits) h.store(0);
    std::atomic<long long> sum(0);
    pool.run(n, w, [&](int i) {
      hits[i].fetch_add(1);
      long long s = 0;
      for (int k = 0; k < (i % 7) * 50; ++k) s += k;  // uneven work
      sum.fetch_add(i + (s & 0));
    });
    for (int i = 0; i < n; ++i)
      if (hits[i].load() != 1) { fprintf(stderr, "region %d: item %d ran %d times\n", rep, i, hits[i].load()); return 1; }
    if (sum.load() != (long long)n * (n - 1) / 2) { fprintf(stderr, "region %d: wrong sum\n", rep); return 1; }
    ++regions;
  }
  // two threads calling in (regions are serialised inside the pool), as the upload thread and the atlas thread do
  std::atomic<int> bad(0);
  auto caller = [&](int id) {
    for (int rep = 0; rep < 500; ++rep) {
      const int n = 1 + (id * 37 + rep) % 97;
      std::atomic<int> cnt(0);
      pool.run(n, 8, [&](int) { cnt.fetch_add(1); });
      if (cnt.load() != n) bad.fetch_add(1);
    }
  };
  std::thread a(caller, 0), b(caller, 1);
  a.join(); b.join();
  if (bad.load()) { fprintf(stderr, "%d concurrent regions lost items\n", bad.load()); return 1; }
  // a pool that never ran, and one destroyed right after a region
  { HostPool idle; }
  { HostPool p2; p2.run(64, 16, [](int) {}); }
  printf("ok %lld regions, %d worker threads initialised\n", regions + 1000, inits.load());
  return 0;
}
