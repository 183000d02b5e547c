#include "core/trace.h"

#include <sys/stat.h>

#include <chrono>
#include <cstdio>
#include <sstream>

#include "core/env.h"
#include "core/log.h"

namespace bps {

int64_t now_us() {
  return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch())
      .count();
}

Timeline::Timeline() {
  on_ = env_bool("BYTEPS_TRACE_ON", false);
  start_ = (int)env_int("BYTEPS_TRACE_START_STEP", 10);
  end_ = (int)env_int("BYTEPS_TRACE_END_STEP", 20);
  dir_ = env_str("BYTEPS_TRACE_DIR", "./trace");
}

void Timeline::configure(bool on, int start_step, int end_step, const std::string& dir, int local_rank) {
  std::lock_guard<std::mutex> g(mu_);
  on_ = on;
  start_ = start_step;
  end_ = end_step;
  if (!dir.empty()) dir_ = dir;
  local_rank_ = local_rank;
}

void Timeline::record(const std::string& tensor, const std::string& stage, uint64_t key, int64_t ts_us,
                      int64_t dur_us) {
  std::lock_guard<std::mutex> g(mu_);
  events_.push_back(TraceEvent{tensor, stage, key, ts_us, dur_us});
}

size_t Timeline::num_events() const {
  std::lock_guard<std::mutex> g(mu_);
  return events_.size();
}

static std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') {
      o.push_back('\\');
      o.push_back(c);
    } else if (c == '\n') {
      o += "\\n";
    } else {
      o.push_back(c);
    }
  }
  return o;
}

std::string Timeline::to_json() const {
  std::lock_guard<std::mutex> g(mu_);
  std::ostringstream os;
  os << "{\n  \"traceEvents\": [\n";
  bool first = true;
  for (const auto& e : events_) {
    if (!first) os << ",\n";
    first = false;
    std::string comm = "Comm." + json_escape(e.tensor);
    os << "    {\"ph\": \"X\", \"args\": {\"name\": \"" << comm << "\"}, \"pid\": \"" << comm << "\", \"name\": \""
       << comm;
    if (!e.stage.empty()) os << "." << e.stage;
    os << "\", \"ts\": " << e.ts_us << ", \"dur\": " << e.dur_us << ", \"tid\": ";
    if (e.key == ~0ull)
      os << "\"total\"";
    else
      os << "\"" << e.key << "\"";
    os << ", \"cat\": \"Comm\"}";
  }
  os << "\n  ],\n  \"displayTimeUnit\": \"ms\"\n}\n";
  return os.str();
}

std::string Timeline::dump() {
  std::string dir;
  int lr;
  {
    std::lock_guard<std::mutex> g(mu_);
    dir = dir_;
    lr = local_rank_;
  }
  mkdir(dir.c_str(), 0755);
  std::string sub = dir + "/" + std::to_string(lr);
  mkdir(sub.c_str(), 0755);
  std::string path = sub + "/comm.json";
  FILE* f = fopen(path.c_str(), "w");
  if (!f) {
    BPS_LOG(ERROR) << "cannot write trace to " << path;
    return "";
  }
  std::string js = to_json();
  fwrite(js.data(), 1, js.size(), f);
  fclose(f);
  BPS_LOG(INFO) << "timeline written to " << path;
  return path;
}

void Timeline::clear() {
  std::lock_guard<std::mutex> g(mu_);
  events_.clear();
}

void Telemetry::record(size_t bytes) {
  if (!on_) return;
  std::lock_guard<std::mutex> g(mu_);
  int64_t now = now_us();
  if (!init_) {
    init_ = true;
    acc_ = 0;
    last_us_ = now;
  }
  acc_ += bytes;
  total_ += bytes;
  double dt = (now - last_us_) / 1e6;
  if (dt > interval_s_) {
    points_.push_back(SpeedEntry{now / 1000, acc_ / 1.0e6 / dt});
    acc_ = 0;
    last_us_ = now;
    if (points_.size() > 1024) points_.pop_front();
  }
}

SpeedEntry Telemetry::get() {
  std::lock_guard<std::mutex> g(mu_);
  if (points_.empty()) return SpeedEntry{0, -5.0};
  SpeedEntry e = points_.front();
  points_.pop_front();
  return e;
}

}  // namespace bps
