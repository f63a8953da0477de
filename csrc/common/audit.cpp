#include "common/audit.h"

#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common/log.h"

namespace bb::audit {
namespace {
std::mutex g_mu;
int g_fd = -1;
std::string g_path;
bool g_init = false;
std::atomic<bool> g_on{false};
std::atomic<uint64_t> g_written{0};
thread_local std::string_view t_who, t_peer;

void append_escaped(std::string& out, std::string_view v) {
  if (v.size() > 256) v = v.substr(0, 256);
  for (const unsigned char ch : v) {
    switch (ch) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      default:
        if (ch < 0x20 || ch == 0x7f) {
          char buf[8];
          std::snprintf(buf, sizeof buf, "\\u%04x", ch);
          out += buf;
        } else {
          out += static_cast<char>(ch);  // bytes >= 0x80 pass through (keys are byte strings; readers decode leniently)
        }
    }
  }
}

void lazy_init_locked() {
  if (g_init) return;
  g_init = true;
  if (const char* e = std::getenv("BB_AUDIT_LOG"); e && *e) {
    const int fd = ::open(e, O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0640);
    if (fd >= 0) {
      g_fd = fd;
      g_path = e;
      g_on.store(true, std::memory_order_release);
    } else {
      BB_LOG(ERROR) << "audit: cannot open " << e << " for appending";
    }
  }
}
}  // namespace

bool open(const std::string& path) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_init = true;
  int fd = -1;
  if (!path.empty()) {
    fd = ::open(path.c_str(), O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0640);
    if (fd < 0) return false;
  }
  if (g_fd >= 0) ::close(g_fd);
  g_fd = fd;
  g_path = path;
  g_on.store(fd >= 0, std::memory_order_release);
  return true;
}

bool enabled() {
  if (g_on.load(std::memory_order_acquire)) return true;
  std::lock_guard<std::mutex> lk(g_mu);
  lazy_init_locked();
  return g_fd >= 0;
}

std::string path() {
  std::lock_guard<std::mutex> lk(g_mu);
  lazy_init_locked();
  return g_path;
}

uint64_t events_written() { return g_written.load(std::memory_order_relaxed); }

void event(std::string_view kind, std::initializer_list<Field> fields) {
  if (!enabled()) return;
  const auto now = std::chrono::system_clock::now().time_since_epoch();
  const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(now).count();
  std::string line;
  line.reserve(192);
  line += "{\"ts\":";
  line += std::to_string(ms / 1000) + "." + (ms % 1000 < 100 ? (ms % 1000 < 10 ? "00" : "0") : "") + std::to_string(ms % 1000);
  line += ",\"event\":\"";
  append_escaped(line, kind);
  line += '"';
  bool has_who = false, has_peer = false;
  for (const auto& f : fields) {
    has_who |= f.first == "who";
    has_peer |= f.first == "peer";
  }
  auto add = [&](std::string_view k, std::string_view v) {
    line += ",\"";
    append_escaped(line, k);
    line += "\":\"";
    append_escaped(line, v);
    line += '"';
  };
  if (!has_who) add("who", t_who.empty() ? std::string_view("local") : t_who);
  if (!has_peer && !t_peer.empty()) add("peer", t_peer);
  for (const auto& f : fields) add(f.first, f.second);
  line += "}\n";
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_fd < 0) return;
  if (::write(g_fd, line.data(), line.size()) == static_cast<ssize_t>(line.size())) g_written.fetch_add(1, std::memory_order_relaxed);
}

Scope::Scope(std::string_view who, std::string_view peer) : prev_who_(t_who), prev_peer_(t_peer) {
  t_who = who;
  t_peer = peer;
}
Scope::~Scope() {
  t_who = prev_who_;
  t_peer = prev_peer_;
}

}  // namespace bb::audit
