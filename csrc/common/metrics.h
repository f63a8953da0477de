// Tiny Prometheus-style metrics registry (counters, gauges, histograms) with text exposition.
// The reference opens a metrics port but registers no route (rpc_service.cpp:212-226).
#pragma once
#include <atomic>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string_view>
#include <string>
#include <vector>

namespace bb {

class Histogram {
 public:
  // bucket upper bounds (inclusive); +Inf is implicit
  explicit Histogram(std::vector<double> bounds);
  void observe(double v);
  uint64_t count() const { return count_.load(); }
  double sum() const;
  double quantile(double q) const;  // bucket-interpolated estimate
  const std::vector<double>& bounds() const { return bounds_; }
  std::vector<uint64_t> bucket_counts() const;

 private:
  std::vector<double> bounds_;
  std::unique_ptr<std::atomic<uint64_t>[]> buckets_;
  std::atomic<uint64_t> count_{0};
  std::atomic<uint64_t> sum_micro_{0};  // sum * 1e3 to keep integer atomics
};

class Metrics {
 public:
  // Hot-path calls take a string_view (no allocation) and only a shared lock once the series exists.
  void inc(std::string_view name, uint64_t by = 1);
  void set_gauge(const std::string& name, double v);
  void observe(std::string_view name, double v);  // latency histograms in microseconds
  // Stable handles for series updated on a hot path (valid for the lifetime of this Metrics).
  std::atomic<uint64_t>* counter_ref(const std::string& name);
  Histogram* histogram_ref(const std::string& name);
  uint64_t counter(const std::string& name) const;
  double gauge(const std::string& name) const;
  void describe(const std::string& name, const std::string& help);
  // Prometheus text format, metric names prefixed with `prefix`.
  std::string render(const std::string& prefix = "bb_") const;
  static std::vector<double> default_latency_bounds_us();
  // name -> (count, sum, p50, p99) of every histogram
  std::map<std::string, std::vector<double>> histogram_summary() const;

 private:
  mutable std::shared_mutex mu_;
  std::map<std::string, std::unique_ptr<std::atomic<uint64_t>>, std::less<>> counters_;
  std::map<std::string, double> gauges_;
  std::map<std::string, std::unique_ptr<Histogram>, std::less<>> hists_;
  std::map<std::string, std::string> help_;
};

}  // namespace bb
