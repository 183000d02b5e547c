// Communication timeline (Chrome trace) + push-pull speed telemetry.
//
// Parity: BYTEPS_TRACE_ON/_START_STEP/_END_STEP/_DIR handling and JSON layout in
// /root/reference/byteps/common/global.cc:448-564 (docs/timeline.md:33-72) and
// PushPullSpeed in global.cc:697-752.  Unlike the reference the recorder is
// thread-safe, and spans may be supplied with device-measured durations.
#pragma once
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

namespace bps {

int64_t now_us();

struct TraceEvent {
  std::string tensor;   // user-visible tensor name
  std::string stage;    // "" for the whole-tensor span
  uint64_t key;         // partition key (tid); ~0 for total
  int64_t ts_us;
  int64_t dur_us;
};

class Timeline {
 public:
  Timeline();
  void configure(bool on, int start_step, int end_step, const std::string& dir, int local_rank);
  bool enabled() const { return on_; }
  // step-window gate: record only when start_step <= step < end_step
  bool active(int64_t step) const { return on_ && step >= start_ && step < end_; }
  int start_step() const { return start_; }
  int end_step() const { return end_; }
  void record(const std::string& tensor, const std::string& stage, uint64_t key, int64_t ts_us, int64_t dur_us);
  size_t num_events() const;
  std::string to_json() const;
  // Writes <dir>/<local_rank>/comm.json; returns the path ("" on failure).
  std::string dump();
  void clear();

 private:
  bool on_ = false;
  int start_ = 10, end_ = 20;
  std::string dir_ = "./trace";
  int local_rank_ = 0;
  mutable std::mutex mu_;
  std::vector<TraceEvent> events_;
};

struct SpeedEntry {
  int64_t ts_ms;
  double mbps;
};

class Telemetry {
 public:
  explicit Telemetry(bool on = true, double interval_s = 10.0) : on_(on), interval_s_(interval_s) {}
  void configure(bool on, double interval_s) { on_ = on; interval_s_ = interval_s; }
  void record(size_t bytes);
  // Pops the oldest data point; {0, -5.0} when none (same sentinel as the reference).
  SpeedEntry get();
  bool should_record() const { return on_; }
  uint64_t total_bytes() const { return total_; }

 private:
  bool on_;
  double interval_s_;
  std::mutex mu_;
  bool init_ = false;
  int64_t last_us_ = 0;
  size_t acc_ = 0;
  uint64_t total_ = 0;
  std::deque<SpeedEntry> points_;
};

}  // namespace bps
