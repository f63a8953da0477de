// Minimal levelled logger (the reference uses glog everywhere; glog is unavailable offline).
// Usage: BB_LOG(INFO) << "x=" << x;   BB_VLOG(2) << ...;   level from BB_LOG_LEVEL / set_log_level().
#pragma once
#include <sstream>
#include <string>

namespace bb {

enum class LogLevel : int { DEBUG = 0, INFO = 1, WARNING = 2, ERROR = 3, OFF = 4 };

void set_log_level(LogLevel l) noexcept;
LogLevel log_level() noexcept;
void set_log_verbosity(int v) noexcept;  // for BB_VLOG(n)
int log_verbosity() noexcept;
void set_log_file(const std::string& path);  // "" = stderr

class LogMessage {
 public:
  LogMessage(LogLevel l, const char* file, int line);
  ~LogMessage();
  std::ostringstream& stream() { return ss_; }

 private:
  LogLevel level_;
  std::ostringstream ss_;
};

struct LogVoidify {
  void operator&(std::ostream&) {}
};

}  // namespace bb

#define BB_LOG_ENABLED(lvl) (static_cast<int>(::bb::LogLevel::lvl) >= static_cast<int>(::bb::log_level()))
#define BB_LOG(lvl) \
  !BB_LOG_ENABLED(lvl) ? (void)0 : ::bb::LogVoidify() & ::bb::LogMessage(::bb::LogLevel::lvl, __FILE__, __LINE__).stream()
#define BB_VLOG(n) \
  !(::bb::log_verbosity() >= (n)) ? (void)0 : ::bb::LogVoidify() & ::bb::LogMessage(::bb::LogLevel::DEBUG, __FILE__, __LINE__).stream()
