// Leveled logging + fatal checks.  Parity: /root/reference/byteps/common/logging.h:26-102
// (BPS_LOG / BPS_CHECK, level from BYTEPS_LOG_LEVEL, BYTEPS_LOG_HIDE_TIME).
#pragma once
#include <sstream>
#include <string>

namespace bps {

enum LogLevel : int { L_TRACE = 0, L_DEBUG = 1, L_INFO = 2, L_WARNING = 3, L_ERROR = 4, L_FATAL = 5 };

int min_log_level();            // parsed once from BYTEPS_LOG_LEVEL (default WARNING)
void set_min_log_level(int l);  // used by tests
bool log_hide_time();

class LogMessage {
 public:
  LogMessage(const char* file, int line, int level);
  ~LogMessage() noexcept(false);
  std::ostringstream& stream() { return ss_; }

 private:
  std::ostringstream ss_;
  const char* file_;
  int line_;
  int level_;
};

struct LogVoidify {
  void operator&(std::ostream&) {}
};

}  // namespace bps

#define BPS_LOG_AT(lvl) \
  (lvl) < ::bps::min_log_level() ? (void)0 : ::bps::LogVoidify() & ::bps::LogMessage(__FILE__, __LINE__, (lvl)).stream()
#define BPS_LOG_TRACE BPS_LOG_AT(::bps::L_TRACE)
#define BPS_LOG_DEBUG BPS_LOG_AT(::bps::L_DEBUG)
#define BPS_LOG_INFO BPS_LOG_AT(::bps::L_INFO)
#define BPS_LOG_WARNING BPS_LOG_AT(::bps::L_WARNING)
#define BPS_LOG_ERROR BPS_LOG_AT(::bps::L_ERROR)
#define BPS_LOG_FATAL ::bps::LogMessage(__FILE__, __LINE__, ::bps::L_FATAL).stream()
#define BPS_LOG(sev) BPS_LOG_##sev

#define BPS_CHECK(cond) \
  if (!(cond)) BPS_LOG_FATAL << "Check failed: " #cond " "
#define BPS_CHECK_OP(a, b, op) \
  if (!((a)op(b))) BPS_LOG_FATAL << "Check failed: " #a " " #op " " #b " (" << (a) << " vs " << (b) << ") "
#define BPS_CHECK_EQ(a, b) BPS_CHECK_OP(a, b, ==)
#define BPS_CHECK_NE(a, b) BPS_CHECK_OP(a, b, !=)
#define BPS_CHECK_LE(a, b) BPS_CHECK_OP(a, b, <=)
#define BPS_CHECK_LT(a, b) BPS_CHECK_OP(a, b, <)
#define BPS_CHECK_GE(a, b) BPS_CHECK_OP(a, b, >=)
#define BPS_CHECK_GT(a, b) BPS_CHECK_OP(a, b, >)
