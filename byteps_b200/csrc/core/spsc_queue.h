// Bounded single-producer/single-consumer ring with busy polling.
//
// Parity: ps-lite's lockless receive queue (/root/reference/3rdparty/ps-lite/include/ps/
// internal/spsc_queue.h + threadsafe_queue.h, enabled with DMLC_LOCKLESS_QUEUE): the van's
// receiving thread hands messages to a customer thread without a mutex/condvar wake-up on
// the critical path.  Producers are serialised by a test-and-set flag so a second producer
// (loopback sends) stays correct; the consumer side is wait-free.
#pragma once

#include <atomic>
#include <chrono>
#include <cstddef>
#include <thread>
#include <utility>
#include <vector>

namespace bps {

template <typename T>
class SpscQueue {
 public:
  explicit SpscQueue(size_t capacity_pow2 = 4096) : buf_(round_up(capacity_pow2)), mask_(buf_.size() - 1) {}

  // blocks (spinning, then yielding) while the ring is full
  void push(T v) {
    while (producer_busy_.test_and_set(std::memory_order_acquire)) cpu_relax();
    const size_t t = tail_.load(std::memory_order_relaxed);
    size_t spins = 0;
    while (t - head_.load(std::memory_order_acquire) >= buf_.size()) backoff(spins++);
    buf_[t & mask_] = std::move(v);
    tail_.store(t + 1, std::memory_order_release);
    producer_busy_.clear(std::memory_order_release);
  }

  bool try_pop(T* out) {
    const size_t h = head_.load(std::memory_order_relaxed);
    if (h == tail_.load(std::memory_order_acquire)) return false;
    *out = std::move(buf_[h & mask_]);
    head_.store(h + 1, std::memory_order_release);
    return true;
  }

  // busy-poll until an element arrives or `stop` becomes true (then drains what is left)
  bool wait_pop(T* out, const std::atomic<bool>& stop) {
    size_t spins = 0;
    while (true) {
      if (try_pop(out)) return true;
      if (stop.load(std::memory_order_acquire)) return try_pop(out);
      backoff(spins++);
    }
  }

  size_t size() const { return tail_.load(std::memory_order_acquire) - head_.load(std::memory_order_acquire); }
  bool empty() const { return size() == 0; }

 private:
  static size_t round_up(size_t n) {
    size_t p = 2;
    while (p < n) p <<= 1;
    return p;
  }
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  static void backoff(size_t spins) {
    if (spins < 2000) cpu_relax();
    else if (spins < 4000) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(50));
  }

  std::vector<T> buf_;
  const size_t mask_;
  alignas(64) std::atomic<size_t> head_{0};
  alignas(64) std::atomic<size_t> tail_{0};
  alignas(64) std::atomic_flag producer_busy_ = ATOMIC_FLAG_INIT;
};

}  // namespace bps
