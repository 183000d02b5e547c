// key -> arrival count barrier.  Parity: /root/reference/byteps/common/ready_table.cc:24-44.
// (The reference's SetReadyCount lacks a return statement; ours is void.)
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <unordered_map>

namespace bps {

class ReadyTable {
 public:
  ReadyTable(int ready_count, std::string name) : ready_count_(ready_count), name_(std::move(name)) {}
  bool is_key_ready(uint64_t key) const {
    std::lock_guard<std::mutex> g(mu_);
    auto it = table_.find(key);
    return it != table_.end() && it->second >= ready_count_;
  }
  int add_ready_count(uint64_t key) {
    int v;
    {
      std::lock_guard<std::mutex> g(mu_);
      v = ++table_[key];
    }
    cv_.notify_all();
    return v;
  }
  void set_ready_count(uint64_t key, int cnt) {
    {
      std::lock_guard<std::mutex> g(mu_);
      table_[key] = cnt;
    }
    cv_.notify_all();
  }
  // block until the key has reached the ready count (timeout_ms < 0: forever); the reference polls
  // IsKeyReady from its stage loops, a waiter here sleeps on a condition variable instead
  bool wait_ready(uint64_t key, int64_t timeout_ms = -1) const {
    std::unique_lock<std::mutex> lk(mu_);
    auto ready = [&] {
      auto it = table_.find(key);
      return it != table_.end() && it->second >= ready_count_;
    };
    if (timeout_ms < 0) {
      cv_.wait(lk, ready);
      return true;
    }
    return cv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), ready);
  }
  void clear_ready_count(uint64_t key) {
    std::lock_guard<std::mutex> g(mu_);
    table_.erase(key);
  }
  int count(uint64_t key) const {
    std::lock_guard<std::mutex> g(mu_);
    auto it = table_.find(key);
    return it == table_.end() ? 0 : it->second;
  }
  int ready_count() const { return ready_count_; }
  const std::string& name() const { return name_; }

 private:
  mutable std::mutex mu_;
  mutable std::condition_variable cv_;
  std::unordered_map<uint64_t, int> table_;
  int ready_count_;
  std::string name_;
};

}  // namespace bps
