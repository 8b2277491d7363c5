// Stand-in for <glog/logging.h> so the reference's header-only path (mcts/common.h:10)
// compiles without glog. TEST INFRASTRUCTURE ONLY (oracle/_ref build) - not product code.
// Every stream is swallowed; FATAL conditions terminate like glog would.
#ifndef MAMCTS_ORACLE_SHIM_GLOG_H_
#define MAMCTS_ORACLE_SHIM_GLOG_H_
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>

namespace oracle_shim {
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct FatalStream {
  bool armed;
  std::ostringstream os;
  explicit FatalStream(bool a) : armed(a) {}
  template <class T> FatalStream& operator<<(const T& v) { if (armed) os << v; return *this; }
  FatalStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
  ~FatalStream() { if (armed) { std::cerr << "FATAL: " << os.str() << std::endl; std::abort(); } }
};
}  // namespace oracle_shim

#define ORACLE_SHIM_SEV_INFO    ::oracle_shim::NullStream()
#define ORACLE_SHIM_SEV_WARNING ::oracle_shim::NullStream()
#define ORACLE_SHIM_SEV_ERROR   ::oracle_shim::NullStream()
#define ORACLE_SHIM_SEV_FATAL   ::oracle_shim::FatalStream(true)
#define LOG(sev) ORACLE_SHIM_SEV_##sev
#define ORACLE_SHIM_IF_INFO(c)    ::oracle_shim::NullStream()
#define ORACLE_SHIM_IF_WARNING(c) ::oracle_shim::NullStream()
#define ORACLE_SHIM_IF_ERROR(c)   ::oracle_shim::NullStream()
#define ORACLE_SHIM_IF_FATAL(c)   ::oracle_shim::FatalStream(static_cast<bool>(c))
#define LOG_IF(sev, cond) ORACLE_SHIM_IF_##sev(cond)
#define LOG_EVERY_N(sev, n) ::oracle_shim::NullStream()
#define LOG_FIRST_N(sev, n) ::oracle_shim::NullStream()
#define VLOG(l) ::oracle_shim::NullStream()
#define VLOG_IF(l, c) ::oracle_shim::NullStream()
#define VLOG_EVERY_N(l, n) ::oracle_shim::NullStream()
#define DLOG(sev) ::oracle_shim::NullStream()
#define DVLOG(l) ::oracle_shim::NullStream()
#define CHECK(c) ORACLE_SHIM_IF_FATAL(!(c))
#endif
