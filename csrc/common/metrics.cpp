#include "common/metrics.h"

#include <algorithm>
#include <cstdio>
#include <sstream>

namespace bb {

Histogram::Histogram(std::vector<double> bounds) : bounds_(std::move(bounds)) {
  std::sort(bounds_.begin(), bounds_.end());
  buckets_.reset(new std::atomic<uint64_t>[bounds_.size() + 1]);
  for (size_t i = 0; i <= bounds_.size(); ++i) buckets_[i].store(0);
}

void Histogram::observe(double v) {
  size_t i = static_cast<size_t>(std::lower_bound(bounds_.begin(), bounds_.end(), v) - bounds_.begin());
  buckets_[i].fetch_add(1, std::memory_order_relaxed);
  count_.fetch_add(1, std::memory_order_relaxed);
  sum_micro_.fetch_add(static_cast<uint64_t>(v < 0 ? 0 : v * 1000.0), std::memory_order_relaxed);
}

double Histogram::sum() const { return static_cast<double>(sum_micro_.load()) / 1000.0; }

std::vector<uint64_t> Histogram::bucket_counts() const {
  std::vector<uint64_t> v(bounds_.size() + 1);
  for (size_t i = 0; i < v.size(); ++i) v[i] = buckets_[i].load();
  return v;
}

double Histogram::quantile(double q) const {
  const auto counts = bucket_counts();
  uint64_t total = 0;
  for (auto c : counts) total += c;
  if (!total) return 0.0;
  const double target = q * static_cast<double>(total);
  double cum = 0;
  for (size_t i = 0; i < counts.size(); ++i) {
    const double next = cum + static_cast<double>(counts[i]);
    if (next >= target && counts[i]) {
      const double lo = i == 0 ? 0.0 : bounds_[i - 1];
      const double hi = i < bounds_.size() ? bounds_[i] : bounds_.back() * 2;
      return lo + (hi - lo) * (target - cum) / static_cast<double>(counts[i]);
    }
    cum = next;
  }
  return bounds_.empty() ? 0.0 : bounds_.back();
}

std::vector<double> Metrics::default_latency_bounds_us() {
  return {1, 2, 5, 10, 20, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000, 50000, 100000, 500000, 1000000};
}

void Metrics::inc(std::string_view name, uint64_t by) {
  std::atomic<uint64_t>* c = nullptr;
  {
    std::shared_lock<std::shared_mutex> lk(mu_);
    auto it = counters_.find(name);
    if (it != counters_.end()) c = it->second.get();
  }
  if (!c) {
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto& slot = counters_[std::string(name)];
    if (!slot) slot = std::make_unique<std::atomic<uint64_t>>(0);
    c = slot.get();
  }
  c->fetch_add(by, std::memory_order_relaxed);
}

void Metrics::set_gauge(const std::string& name, double v) {
  std::unique_lock<std::shared_mutex> lk(mu_);
  gauges_[name] = v;
}

void Metrics::observe(std::string_view name, double v) {
  Histogram* h = nullptr;
  {
    std::shared_lock<std::shared_mutex> lk(mu_);
    auto it = hists_.find(name);
    if (it != hists_.end()) h = it->second.get();
  }
  if (!h) {
    std::unique_lock<std::shared_mutex> lk(mu_);
    auto& slot = hists_[std::string(name)];
    if (!slot) slot = std::make_unique<Histogram>(default_latency_bounds_us());
    h = slot.get();
  }
  h->observe(v);
}

std::atomic<uint64_t>* Metrics::counter_ref(const std::string& name) {
  std::unique_lock<std::shared_mutex> lk(mu_);
  auto& slot = counters_[name];
  if (!slot) slot = std::make_unique<std::atomic<uint64_t>>(0);
  return slot.get();
}

Histogram* Metrics::histogram_ref(const std::string& name) {
  std::unique_lock<std::shared_mutex> lk(mu_);
  auto& slot = hists_[name];
  if (!slot) slot = std::make_unique<Histogram>(default_latency_bounds_us());
  return slot.get();
}

uint64_t Metrics::counter(const std::string& name) const {
  std::unique_lock<std::shared_mutex> lk(mu_);
  auto it = counters_.find(name);
  return it == counters_.end() ? 0 : it->second->load();
}

double Metrics::gauge(const std::string& name) const {
  std::unique_lock<std::shared_mutex> lk(mu_);
  auto it = gauges_.find(name);
  return it == gauges_.end() ? 0.0 : it->second;
}

void Metrics::describe(const std::string& name, const std::string& help) {
  std::unique_lock<std::shared_mutex> lk(mu_);
  help_[name] = help;
}

std::map<std::string, std::vector<double>> Metrics::histogram_summary() const {
  std::map<std::string, std::vector<double>> out;
  std::unique_lock<std::shared_mutex> lk(mu_);
  for (const auto& [n, h] : hists_) out[n] = {static_cast<double>(h->count()), h->sum(), h->quantile(0.5), h->quantile(0.99)};
  return out;
}

std::string Metrics::render(const std::string& prefix) const {
  std::ostringstream out;
  std::unique_lock<std::shared_mutex> lk(mu_);
  auto head = [&](const std::string& n, const char* type) {
    auto h = help_.find(n);
    if (h != help_.end()) out << "# HELP " << prefix << n << ' ' << h->second << '\n';
    out << "# TYPE " << prefix << n << ' ' << type << '\n';
  };
  for (const auto& [n, c] : counters_) {
    head(n, "counter");
    out << prefix << n << ' ' << c->load() << '\n';
  }
  for (const auto& [n, g] : gauges_) {
    head(n, "gauge");
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.10g", g);
    out << prefix << n << ' ' << buf << '\n';
  }
  for (const auto& [n, h] : hists_) {
    head(n, "histogram");
    const auto counts = h->bucket_counts();
    uint64_t cum = 0;
    for (size_t i = 0; i < h->bounds().size(); ++i) {
      cum += counts[i];
      char buf[64];
      std::snprintf(buf, sizeof buf, "%g", h->bounds()[i]);
      out << prefix << n << "_bucket{le=\"" << buf << "\"} " << cum << '\n';
    }
    cum += counts.back();
    out << prefix << n << "_bucket{le=\"+Inf\"} " << cum << '\n';
    char sb[64];
    std::snprintf(sb, sizeof sb, "%.6g", h->sum());
    out << prefix << n << "_sum " << sb << '\n';
    out << prefix << n << "_count " << h->count() << '\n';
  }
  return out.str();
}

}  // namespace bb
