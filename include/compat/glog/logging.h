// Stand-in for <glog/logging.h>: LOG(INFO)/LOG_IF to stderr, DLOG_ASSERT as a
// no-op stream (the reference only ever evaluates it in debug builds).
#pragma once
#include <iostream>
#include <sstream>
#include <string>
extern "C++" {
inline int& scalesim_compat_minloglevel() { static int v = 0; return v; }
inline int& scalesim_compat_logtostderr() { static int v = 1; return v; }
}
#define FLAGS_minloglevel scalesim_compat_minloglevel()
#define FLAGS_logtostderr scalesim_compat_logtostderr()
namespace google {
enum { GLOG_INFO = 0, GLOG_WARNING = 1, GLOG_ERROR = 2, GLOG_FATAL = 3 };
class LogMessage {
 public:
  LogMessage(const char* file, int line, int sev, bool on) : sev_(sev), on_(on) {
    if (on_) {
      const char* base = file;
      for (const char* p = file; *p; ++p) if (*p == '/') base = p + 1;
      static const char tag[] = {'I', 'W', 'E', 'F'};
      ss_ << tag[sev & 3] << ' ' << base << ':' << line << "] ";
    }
  }
  ~LogMessage() {
    if (on_ && sev_ >= FLAGS_minloglevel) { ss_ << '\n'; std::cerr << ss_.str(); std::cerr.flush(); }
    if (on_ && sev_ == GLOG_FATAL) std::abort();
  }
  std::ostream& stream() { return ss_; }
 private:
  std::ostringstream ss_;
  int sev_;
  bool on_;
};
struct NullStream {
  template <class T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
inline void ShutdownGoogleLogging() {}
}  // namespace google
#define SCALESIM_COMPAT_SEV_INFO ::google::GLOG_INFO
#define SCALESIM_COMPAT_SEV_WARNING ::google::GLOG_WARNING
#define SCALESIM_COMPAT_SEV_ERROR ::google::GLOG_ERROR
#define SCALESIM_COMPAT_SEV_FATAL ::google::GLOG_FATAL
#define LOG(sev) ::google::LogMessage(__FILE__, __LINE__, SCALESIM_COMPAT_SEV_##sev, true).stream()
#define LOG_IF(sev, cond) ::google::LogMessage(__FILE__, __LINE__, SCALESIM_COMPAT_SEV_##sev, static_cast<bool>(cond)).stream()
#define DLOG(sev) if (true) {} else ::google::NullStream()
#define DLOG_IF(sev, cond) if (true) {} else ::google::NullStream()
#define DLOG_ASSERT(cond) if (true) {} else ::google::NullStream()
#define CHECK(cond) LOG_IF(FATAL, !(cond)) << "Check failed: " #cond " "
