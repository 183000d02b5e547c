// Fixed-size worker pool for compress/decompress and host reductions.
// Parity: /root/reference/byteps/common/thread_pool.h:17-75.  Unlike the
// reference (pool only exists when BYTEPS_THREADPOOL_SIZE is set, yet COMPRESS
// dereferences it) the pool is always constructed with a sane default.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace bps {

class ThreadPool {
 public:
  explicit ThreadPool(size_t n) {
    if (n == 0) n = 1;
    for (size_t i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~ThreadPool() { shutdown(); }
  void enqueue(std::function<void()> fn) {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stop_) return;
      q_.push(std::move(fn));
    }
    cv_.notify_one();
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stop_) return;
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_)
      if (t.joinable()) t.join();
  }
  size_t size() const { return workers_.size(); }

 private:
  void loop() {
    for (;;) {
      std::function<void()> fn;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
        if (stop_ && q_.empty()) return;
        fn = std::move(q_.front());
        q_.pop();
      }
      fn();
    }
  }
  std::vector<std::thread> workers_;
  std::queue<std::function<void()>> q_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false;
};

}  // namespace bps
