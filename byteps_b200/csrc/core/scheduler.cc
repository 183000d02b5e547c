#include "core/scheduler.h"

#include <chrono>

namespace bps {

static constexpr uint64_t kUnlimited = 1ull << 62;

ScheduledQueue::ScheduledQueue(int stage, bool scheduled, uint64_t credits_bytes, ReadyTable* rt)
    : stage_(stage),
      scheduled_(scheduled),
      credit_cap_(credits_bytes ? credits_bytes : kUnlimited),
      credits_(credit_cap_),
      rt_(rt),
      q_(Order{scheduled}) {}

void ScheduledQueue::add(const TaskPtr& t) {
  {
    std::lock_guard<std::mutex> g(mu_);
    t->seq = seq_++;
    q_.insert(t);
  }
  cv_.notify_one();
}

TaskPtr ScheduledQueue::get_locked() {
  for (auto it = q_.begin(); it != q_.end(); ++it) {
    const TaskPtr& t = *it;
    if (t->ready && !t->ready()) continue;
    // a task larger than the whole window may still run when the window is idle
    if (t->len > credits_ && credits_ != credit_cap_) continue;
    if (rt_ && !rt_->is_key_ready(t->key)) continue;
    TaskPtr out = t;
    q_.erase(it);
    credits_ -= (out->len > credits_) ? credits_ : out->len;
    if (rt_) rt_->clear_ready_count(out->key);
    return out;
  }
  return nullptr;
}

TaskPtr ScheduledQueue::get() {
  std::lock_guard<std::mutex> g(mu_);
  return get_locked();
}

TaskPtr ScheduledQueue::wait_get(int64_t timeout_us) {
  std::unique_lock<std::mutex> lk(mu_);
  auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(timeout_us);
  while (!stopped_) {
    TaskPtr t = get_locked();
    if (t) return t;
    // Readiness predicates (device events) are polled: wake at least every 50us
    // while something is queued, otherwise sleep until notified.
    auto wake = q_.empty() ? deadline : std::min(deadline, std::chrono::steady_clock::now() + std::chrono::microseconds(50));
    if (cv_.wait_until(lk, wake) == std::cv_status::timeout && std::chrono::steady_clock::now() >= deadline) {
      return get_locked();
    }
  }
  return nullptr;
}

TaskPtr ScheduledQueue::get_by_key(uint64_t key) {
  std::lock_guard<std::mutex> g(mu_);
  for (auto it = q_.begin(); it != q_.end(); ++it) {
    if ((*it)->key == key) {
      if ((*it)->ready && !(*it)->ready()) return nullptr;
      TaskPtr out = *it;
      q_.erase(it);
      return out;
    }
  }
  return nullptr;
}

void ScheduledQueue::report_finish(uint64_t bytes) {
  {
    std::lock_guard<std::mutex> g(mu_);
    credits_ += bytes;
    if (credits_ > credit_cap_) credits_ = credit_cap_;
  }
  cv_.notify_all();
}

void ScheduledQueue::notify() { cv_.notify_all(); }

void ScheduledQueue::stop() {
  {
    std::lock_guard<std::mutex> g(mu_);
    stopped_ = true;
  }
  cv_.notify_all();
}

size_t ScheduledQueue::pending() const {
  std::lock_guard<std::mutex> g(mu_);
  return q_.size();
}

uint64_t ScheduledQueue::credits() const {
  std::lock_guard<std::mutex> g(mu_);
  return credits_;
}

void ScheduledQueue::reset() {
  std::lock_guard<std::mutex> g(mu_);
  q_.clear();
  credits_ = credit_cap_;
  stopped_ = false;
  seq_ = 0;
}

}  // namespace bps
