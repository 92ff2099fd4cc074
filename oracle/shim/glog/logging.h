// TEST INFRASTRUCTURE (oracle shim) -- not part of the product path.
// No-op stand-in for glog so the reference gridworld sources compile.
#ifndef ORACLE_SHIM_GLOG_LOGGING_H_
#define ORACLE_SHIM_GLOG_LOGGING_H_

#include <iostream>

namespace oracle_shim {
struct NullStream {
  template <typename T>
  NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace oracle_shim

#define VLOG(n) ::oracle_shim::NullStream()
#define LOG(sev) ::oracle_shim::NullStream()
#define LOG_IF(sev, cond) ::oracle_shim::NullStream()

#endif  // ORACLE_SHIM_GLOG_LOGGING_H_
