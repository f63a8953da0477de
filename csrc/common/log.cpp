#include "common/log.h"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>

#include <sys/syscall.h>
#include <unistd.h>

namespace bb {
namespace {
std::atomic<int> g_level{-1};
std::atomic<int> g_verbosity{-1};
std::mutex g_mu;
FILE* g_file = nullptr;

int init_level() {
  const char* e = std::getenv("BB_LOG_LEVEL");
  int lvl = static_cast<int>(LogLevel::WARNING);
  if (e) {
    if (!strcasecmp(e, "debug")) lvl = 0;
    else if (!strcasecmp(e, "info")) lvl = 1;
    else if (!strcasecmp(e, "warning") || !strcasecmp(e, "warn")) lvl = 2;
    else if (!strcasecmp(e, "error")) lvl = 3;
    else if (!strcasecmp(e, "off")) lvl = 4;
  }
  return lvl;
}
}  // namespace

void set_log_level(LogLevel l) noexcept { g_level.store(static_cast<int>(l)); }
LogLevel log_level() noexcept {
  int l = g_level.load(std::memory_order_relaxed);
  if (l < 0) {
    l = init_level();
    g_level.store(l);
  }
  return static_cast<LogLevel>(l);
}
void set_log_verbosity(int v) noexcept { g_verbosity.store(v); }
int log_verbosity() noexcept {
  int v = g_verbosity.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = std::getenv("BB_VLOG");
    v = e ? std::atoi(e) : 0;
    g_verbosity.store(v);
  }
  return v;
}
void set_log_file(const std::string& path) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_file) {
    std::fclose(g_file);
    g_file = nullptr;
  }
  if (!path.empty()) g_file = std::fopen(path.c_str(), "a");
}

LogMessage::LogMessage(LogLevel l, const char* file, int line) : level_(l) {
  static const char kTag[] = {'D', 'I', 'W', 'E'};
  auto now = std::chrono::system_clock::now();
  auto t = std::chrono::system_clock::to_time_t(now);
  auto us = std::chrono::duration_cast<std::chrono::microseconds>(now.time_since_epoch()).count() % 1000000;
  struct tm tmv;
  localtime_r(&t, &tmv);
  char buf[64];
  std::snprintf(buf, sizeof buf, "%c%02d%02d %02d:%02d:%02d.%06ld ", kTag[static_cast<int>(l) & 3], tmv.tm_mon + 1,
                tmv.tm_mday, tmv.tm_hour, tmv.tm_min, tmv.tm_sec, static_cast<long>(us));
  const char* base = std::strrchr(file, '/');
  ss_ << buf << static_cast<long>(::syscall(SYS_gettid)) << ' ' << (base ? base + 1 : file) << ':' << line << "] ";
}

LogMessage::~LogMessage() {
  ss_ << '\n';
  const std::string s = ss_.str();
  std::lock_guard<std::mutex> lk(g_mu);
  FILE* f = g_file ? g_file : stderr;
  std::fwrite(s.data(), 1, s.size(), f);
  if (level_ >= LogLevel::WARNING || g_file) std::fflush(f);
}

}  // namespace bb
