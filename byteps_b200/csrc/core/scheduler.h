// Priority + byte-credit scheduler, one instance per pipeline stage.
//
// Parity: BytePSScheduledQueue (/root/reference/byteps/common/scheduled_queue.cc:26-210).
// Same admission rule - a task is dispatched when (a) its ready predicate holds,
// (b) it fits in the remaining byte credits and (c) its key is ready in the
// stage's ReadyTable - but tasks live in an ordered set keyed
// (priority desc, key asc, arrival) instead of a vector that is re-sorted on
// every insert, and consumers block on a condition variable instead of
// spinning with 1us sleeps.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <memory>
#include <mutex>
#include <set>
#include <unordered_map>

#include "core/ready_table.h"
#include "core/task.h"

namespace bps {

class ScheduledQueue {
 public:
  // credits_bytes == 0 -> unlimited.  scheduled == false -> FIFO order.
  ScheduledQueue(int stage, bool scheduled, uint64_t credits_bytes, ReadyTable* rt = nullptr);

  void add(const TaskPtr& t);
  // Non-blocking: returns nullptr when nothing is eligible.
  TaskPtr get();
  // Blocking with timeout (microseconds); returns nullptr on timeout/stop.
  TaskPtr wait_get(int64_t timeout_us);
  // Take the task with this exact key (follower ranks obeying the root's order).
  TaskPtr get_by_key(uint64_t key);
  void report_finish(uint64_t bytes);
  void notify();  // external readiness changed (event fired, table updated)
  void stop();
  size_t pending() const;
  uint64_t credits() const;
  int stage() const { return stage_; }
  bool scheduled() const { return scheduled_; }
  void reset();

 private:
  struct Order {
    bool scheduled;
    bool operator()(const TaskPtr& a, const TaskPtr& b) const {
      if (scheduled) {
        if (a->priority != b->priority) return a->priority > b->priority;
        if (a->key != b->key) return a->key < b->key;
      }
      return a->seq < b->seq;
    }
  };
  TaskPtr get_locked();

  int stage_;
  bool scheduled_;
  uint64_t credit_cap_;
  uint64_t credits_;
  ReadyTable* rt_;
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::set<TaskPtr, Order> q_;
  uint64_t seq_ = 0;
  bool stopped_ = false;
};

}  // namespace bps
