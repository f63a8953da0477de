// Tracing (the reference has only ad-hoc chrono spans in its CLIs).
//
//  * TraceSpan: RAII host span.  When tracing is on (BB_TRACE=<file> in the environment, or
//    trace::enable()) spans are recorded into a bounded in-memory ring and written as Chrome /
//    Perfetto "traceEvents" JSON by trace::dump() (automatically at exit when BB_TRACE is set).
//  * Every span is also an NVTX range (nvtx3 is header-only; it is a no-op unless a profiler such as
//    nsys / ncu injects itself), so the same phases show up on a GPU timeline next to the kernels.
//  * Off by default: a disabled span costs one relaxed atomic load.
#pragma once
#include <atomic>
#include <cstdint>
#include <string>

namespace bb::trace {

bool enabled();
void enable(bool on, size_t ring_capacity = 1u << 16);
// Writes the ring as {"traceEvents":[...]} ; returns the number of events written (0 on I/O failure).
size_t dump(const std::string& path);
size_t recorded();
void clear();
// Instant event with an integer argument (e.g. bytes, object count).
void instant(const char* name, uint64_t arg = 0);

class Span {
 public:
  explicit Span(const char* name, uint64_t arg = 0);
  ~Span();
  Span(const Span&) = delete;
  Span& operator=(const Span&) = delete;

 private:
  const char* name_;
  uint64_t arg_;
  uint64_t t0_ns_;
  bool live_;
};

}  // namespace bb::trace

#define BB_TRACE_CONCAT2(a, b) a##b
#define BB_TRACE_CONCAT(a, b) BB_TRACE_CONCAT2(a, b)
#define BB_TRACE_SPAN(name, ...) ::bb::trace::Span BB_TRACE_CONCAT(bb_trace_span_, __LINE__)(name, ##__VA_ARGS__)
