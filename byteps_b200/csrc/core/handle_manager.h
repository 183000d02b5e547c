// Integer handle -> completion status table.
// Parity: /root/reference/byteps/torch/handle_manager.cc:22-52, but waiting
// blocks on a condition variable instead of polling with 1 ms sleeps
// (/root/reference/byteps/torch/ops.cc:129-135).
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <unordered_map>

#include "core/types.h"

namespace bps {

class HandleManager {
 public:
  int allocate() {
    std::lock_guard<std::mutex> g(mu_);
    int h = next_++;
    results_[h] = nullptr;
    return h;
  }
  void mark_done(int h, const Status& s) {
    {
      std::lock_guard<std::mutex> g(mu_);
      results_[h] = std::make_shared<Status>(s);
    }
    cv_.notify_all();
  }
  bool poll(int h) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = results_.find(h);
    if (it == results_.end()) return true;  // unknown/released handles count as complete
    return it->second != nullptr;
  }
  // Blocks until done (timeout_ms < 0: forever).  Returns the status and frees the slot.
  Status wait_and_release(int h, int64_t timeout_ms = -1) {
    std::unique_lock<std::mutex> lk(mu_);
    auto pred = [&] {
      auto it = results_.find(h);
      return it == results_.end() || it->second != nullptr;
    };
    if (timeout_ms < 0) {
      cv_.wait(lk, pred);
    } else if (!cv_.wait_for(lk, std::chrono::milliseconds(timeout_ms), pred)) {
      return Status::InProgress();
    }
    auto it = results_.find(h);
    if (it == results_.end()) return Status::OK();
    Status s = *it->second;
    results_.erase(it);
    return s;
  }
  size_t outstanding() {
    std::lock_guard<std::mutex> g(mu_);
    size_t n = 0;
    for (auto& kv : results_) n += kv.second == nullptr;
    return n;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::unordered_map<int, std::shared_ptr<Status>> results_;
  int next_ = 0;
};

}  // namespace bps
