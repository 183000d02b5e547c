// getenv helpers: the whole knob surface is environment variables, as in the
// reference (/root/reference/docs/env.md).
#pragma once
#include <cstdlib>
#include <string>

namespace bps {

inline std::string env_str(const char* k, const std::string& dflt = "") {
  const char* v = getenv(k);
  return v ? std::string(v) : dflt;
}
inline long long env_int(const char* k, long long dflt) {
  const char* v = getenv(k);
  if (!v || !*v) return dflt;
  return atoll(v);
}
inline double env_float(const char* k, double dflt) {
  const char* v = getenv(k);
  if (!v || !*v) return dflt;
  return atof(v);
}
inline bool env_bool(const char* k, bool dflt = false) {
  const char* v = getenv(k);
  if (!v || !*v) return dflt;
  return atoi(v) != 0 || v[0] == 't' || v[0] == 'T' || v[0] == 'y' || v[0] == 'Y';
}
inline bool env_has(const char* k) { return getenv(k) != nullptr; }

}  // namespace bps
