#include "common/fault.h"

#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>

namespace bb::fault {
namespace {
struct Fault {
  int64_t value = 1;
  int64_t remaining = -1;
};
std::mutex g_mu;
std::map<std::string, Fault> g_faults;
std::atomic<size_t> g_armed{0};  // fast path: nothing armed -> no lock

struct EnvInit {
  EnvInit() {
    if (const char* s = std::getenv("BB_FAULT")) arm_from_spec(s);
  }
} g_env_init;
}  // namespace

bool any_armed() { return g_armed.load(std::memory_order_relaxed) != 0; }

bool fire(const char* name) {
  if (!any_armed()) return false;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_faults.find(name);
  if (it == g_faults.end()) return false;
  if (it->second.remaining > 0 && --it->second.remaining == 0) {
    g_faults.erase(it);
    g_armed.store(g_faults.size(), std::memory_order_relaxed);
  }
  return true;
}

int64_t value(const char* name, int64_t def) {
  if (!any_armed()) return def;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_faults.find(name);
  return it == g_faults.end() ? def : it->second.value;
}

void arm(const std::string& name, int64_t v, int64_t count) {
  if (name.empty() || count == 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  g_faults[name] = Fault{v, count};
  g_armed.store(g_faults.size(), std::memory_order_relaxed);
}

void disarm(const std::string& name) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_faults.erase(name);
  g_armed.store(g_faults.size(), std::memory_order_relaxed);
}

void clear() {
  std::lock_guard<std::mutex> lk(g_mu);
  g_faults.clear();
  g_armed.store(0, std::memory_order_relaxed);
}

size_t arm_from_spec(const std::string& spec) {
  size_t n = 0, pos = 0;
  while (pos <= spec.size()) {
    size_t end = spec.find(',', pos);
    if (end == std::string::npos) end = spec.size();
    std::string tok = spec.substr(pos, end - pos);
    pos = end + 1;
    while (!tok.empty() && tok.front() == ' ') tok.erase(tok.begin());
    while (!tok.empty() && tok.back() == ' ') tok.pop_back();
    if (tok.empty()) continue;
    int64_t v = 1, count = -1;
    const size_t eq = tok.find('='), colon = tok.find(':');
    std::string name = tok.substr(0, std::min(eq, colon));
    if (eq != std::string::npos) v = std::atoll(tok.c_str() + eq + 1);
    if (colon != std::string::npos) count = std::atoll(tok.c_str() + colon + 1);
    arm(name, v, count);
    ++n;
  }
  return n;
}

}  // namespace bb::fault
