#include "common/trace.h"

#include <sys/syscall.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common/spin.h"

#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define BB_HAVE_NVTX 1
#else
#define BB_HAVE_NVTX 0
#endif

namespace bb::trace {
namespace {

struct Event {
  const char* name;
  uint64_t ts_ns;
  uint64_t dur_ns;  // 0 = instant
  uint64_t arg;
  uint32_t tid;
};

std::atomic<bool> g_on{false};
SpinMutex g_mu;
std::vector<Event> g_ring;
size_t g_cap = 0;
size_t g_next = 0;
uint64_t g_total = 0;
std::string g_exit_path;

uint64_t now_ns() {
  return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

uint32_t tid() {
  static thread_local uint32_t t = static_cast<uint32_t>(::syscall(SYS_gettid));
  return t;
}

void record(const Event& e) {
  std::lock_guard<SpinMutex> lk(g_mu);
  if (g_cap == 0) return;
  if (g_ring.size() < g_cap) g_ring.push_back(e);
  else g_ring[g_next % g_cap] = e;
  ++g_next;
  ++g_total;
}

struct EnvInit {
  EnvInit() {
    const char* p = std::getenv("BB_TRACE");
    if (p && *p) {
      g_exit_path = p;
      enable(true);
      std::atexit([] { dump(g_exit_path); });
    }
  }
} g_env_init;

}  // namespace

bool enabled() { return g_on.load(std::memory_order_relaxed); }

void enable(bool on, size_t ring_capacity) {
  {
    std::lock_guard<SpinMutex> lk(g_mu);
    if (on && g_cap != ring_capacity) {
      g_cap = ring_capacity;
      g_ring.clear();
      g_ring.reserve(std::min<size_t>(g_cap, 4096));
      g_next = 0;
    }
  }
  g_on.store(on, std::memory_order_relaxed);
}

size_t recorded() {
  std::lock_guard<SpinMutex> lk(g_mu);
  return g_ring.size();
}

void clear() {
  std::lock_guard<SpinMutex> lk(g_mu);
  g_ring.clear();
  g_next = 0;
}

void instant(const char* name, uint64_t arg) {
#if BB_HAVE_NVTX
  nvtxMarkA(name);
#endif
  if (!enabled()) return;
  record(Event{name, now_ns(), 0, arg, tid()});
}

size_t dump(const std::string& path) {
  std::vector<Event> copy;
  {
    std::lock_guard<SpinMutex> lk(g_mu);
    copy = g_ring;
  }
  FILE* f = std::fopen(path.c_str(), "w");
  if (!f) return 0;
  std::fprintf(f, "{\"displayTimeUnit\":\"ns\",\"traceEvents\":[\n");
  const int pid = static_cast<int>(::getpid());
  for (size_t i = 0; i < copy.size(); ++i) {
    const Event& e = copy[i];
    if (e.dur_ns)
      std::fprintf(f, "{\"name\":\"%s\",\"ph\":\"X\",\"pid\":%d,\"tid\":%u,\"ts\":%.3f,\"dur\":%.3f,\"args\":{\"n\":%llu}}%s\n", e.name, pid, e.tid,
                   static_cast<double>(e.ts_ns) / 1e3, static_cast<double>(e.dur_ns) / 1e3, static_cast<unsigned long long>(e.arg),
                   i + 1 < copy.size() ? "," : "");
    else
      std::fprintf(f, "{\"name\":\"%s\",\"ph\":\"i\",\"s\":\"t\",\"pid\":%d,\"tid\":%u,\"ts\":%.3f,\"args\":{\"n\":%llu}}%s\n", e.name, pid, e.tid,
                   static_cast<double>(e.ts_ns) / 1e3, static_cast<unsigned long long>(e.arg), i + 1 < copy.size() ? "," : "");
  }
  std::fprintf(f, "]}\n");
  std::fclose(f);
  return copy.size();
}

Span::Span(const char* name, uint64_t arg) : name_(name), arg_(arg), t0_ns_(0), live_(enabled()) {
#if BB_HAVE_NVTX
  nvtxRangePushA(name);
#endif
  if (live_) t0_ns_ = now_ns();
}

Span::~Span() {
#if BB_HAVE_NVTX
  nvtxRangePop();
#endif
  if (live_) {
    const uint64_t t1 = now_ns();
    record(Event{name_, t0_ns_, t1 > t0_ns_ ? t1 - t0_ns_ : 1, arg_, tid()});
  }
}

}  // namespace bb::trace
