// host_pool.h -- the host worker pool of fspt_scene_upload (plain C++, no CUDA: unit-tested on the CPU by
// tests/test_host_pool.py).
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

// fspt_scene_upload splits its host work (pre-passes, record building, atlas interleave) into parallel regions of a few
// hundred microseconds each; spawning threads per region cost more than some regions took (measured: ~20 us per
// pthread_create, 12-16 threads, five regions per upload), so the context keeps its workers.  One region at a time;
// the calling thread works too.
class HostPool {
 public:
  // thread_init runs once on every worker (the library binds it to the context's CUDA device)
  explicit HostPool(std::function<void()> thread_init = nullptr) : thread_init_(std::move(thread_init)) {}
  ~HostPool() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  // fn(i) for every i in [0, n_items), on at most max_workers threads; returns when all items are done
  void run(int n_items, int max_workers, const std::function<void(int)>& fn) {
    if (n_items <= 0) return;
    std::lock_guard<std::mutex> region(region_mu_);
    const int helpers = std::max(0, std::min(n_items, max_workers) - 1);
    {
      std::lock_guard<std::mutex> g(mu_);
      while ((int)threads_.size() < helpers) {
        const int idx = (int)threads_.size();
        threads_.emplace_back([this, idx]() { worker(idx); });
      }
      fn_ = &fn; n_items_ = n_items; next_.store(0); wanted_ = helpers; active_ = helpers; ++gen_;
    }
    if (helpers) cv_.notify_all();
    drain();
    std::unique_lock<std::mutex> g(mu_);
    done_cv_.wait(g, [&]() { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain() {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= n_items_) break;
      (*fn_)(i);
    }
  }
  void worker(int idx) {
    if (thread_init_) thread_init_();
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(mu_);
        cv_.wait(g, [&]() { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (idx >= wanted_) continue;
      }
      drain();
      std::lock_guard<std::mutex> g(mu_);
      if (--active_ == 0) done_cv_.notify_all();
    }
  }
  std::function<void()> thread_init_;
  std::mutex region_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void(int)>* fn_ = nullptr;
  int n_items_ = 0, wanted_ = 0, active_ = 0;
  std::atomic<int> next_{0};
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

