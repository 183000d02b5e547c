#include "core/log.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <stdexcept>

namespace bps {

static int parse_level() {
  const char* e = getenv("BYTEPS_LOG_LEVEL");
  if (!e) return L_WARNING;
  std::string s(e);
  for (auto& c : s) c = toupper(c);
  if (s == "TRACE") return L_TRACE;
  if (s == "DEBUG") return L_DEBUG;
  if (s == "INFO") return L_INFO;
  if (s == "WARNING") return L_WARNING;
  if (s == "ERROR") return L_ERROR;
  if (s == "FATAL") return L_FATAL;
  return L_WARNING;
}

static std::atomic<int> g_level{-1};

int min_log_level() {
  int l = g_level.load(std::memory_order_relaxed);
  if (l < 0) {
    l = parse_level();
    g_level.store(l);
  }
  return l;
}
void set_min_log_level(int l) { g_level.store(l); }

bool log_hide_time() {
  static int hide = [] {
    const char* e = getenv("BYTEPS_LOG_HIDE_TIME");
    return (e && atoi(e) != 0) ? 1 : 0;
  }();
  return hide != 0;
}

LogMessage::LogMessage(const char* file, int line, int level) : file_(file), line_(line), level_(level) {}

LogMessage::~LogMessage() noexcept(false) {
  static const char* names = "TDIWEF";
  const char* base = strrchr(file_, '/');
  base = base ? base + 1 : file_;
  char head[96];
  if (log_hide_time()) {
    snprintf(head, sizeof(head), "[%c %s:%d] ", names[level_], base, line_);
  } else {
    auto now = std::chrono::system_clock::now();
    auto us = std::chrono::duration_cast<std::chrono::microseconds>(now.time_since_epoch()).count();
    time_t sec = static_cast<time_t>(us / 1000000);
    struct tm tmv;
    localtime_r(&sec, &tmv);
    char tbuf[32];
    strftime(tbuf, sizeof(tbuf), "%Y-%m-%d %H:%M:%S", &tmv);
    snprintf(head, sizeof(head), "[%s.%06ld: %c %s:%d] ", tbuf, static_cast<long>(us % 1000000), names[level_], base,
             line_);
  }
  std::string msg = std::string(head) + ss_.str() + "\n";
  fputs(msg.c_str(), stderr);
  fflush(stderr);
  if (level_ == L_FATAL) throw std::runtime_error(ss_.str());
}

}  // namespace bps
