// bb-bench: native benchmark client (reference clients/benchmark_client.cpp: per-iteration put +
// get of `--size` bytes, avg latency and MiB/s; examples/benchmark_disk_backends.cpp: reserve /
// commit / free ops/s on a disk backend).
//   bb-bench client  --keystone host:port --size 1048576 --iterations 50 [--replicas 1] [--max-workers 1] [--batch 1]
//   bb-bench backend --class NVME --path /tmp/x --ops 50 --size 4096
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>

#include "apps/cli_util.h"
#include "client/blackbird_client.h"
#include "worker/storage_backend.h"

using namespace bb;
using Clk = std::chrono::steady_clock;

namespace {
double ms(Clk::time_point a, Clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }
double pct(std::vector<double> v, double q) {
  if (v.empty()) return 0;
  std::sort(v.begin(), v.end());
  return v[std::min(v.size() - 1, static_cast<size_t>(q * static_cast<double>(v.size())))];
}
}  // namespace

int main(int argc, char** argv) {
  auto args = bbapp::parse_args(argc, argv);
  const std::string mode = args.positional.empty() ? "client" : args.positional[0];
  if (args.has("help")) {
    std::printf("usage: bb-bench client|backend [options]\n");
    return 0;
  }
  if (mode == "backend") {
    auto sc = parse_storage_class(args.get("class", "NVME"));
    if (!sc) return 2;
    worker::BackendOptions o;
    o.mount_path = args.get("path", "/tmp/bb_bench_backend");
    auto b = worker::create_storage_backend(*sc, static_cast<uint64_t>(args.num("capacity", 100ll << 20)), o);
    if (!b || b->initialize() != ErrorCode::OK) {
      std::fprintf(stderr, "backend init failed\n");
      return 1;
    }
    const int ops = static_cast<int>(args.num("ops", 50));
    const uint64_t size = static_cast<uint64_t>(args.num("size", 4096));
    std::vector<uint8_t> buf(size, 0xA5);
    std::vector<double> t_res, t_wr, t_com, t_free;
    for (int i = 0; i < ops; ++i) {
      auto a = Clk::now();
      auto tok = b->reserve_shard(size);
      auto c = Clk::now();
      if (!tok.ok()) break;
      b->write(tok.value().remote_addr - b->get_base_address(), buf.data(), size);
      auto d = Clk::now();
      b->commit_shard(tok.value());
      auto e = Clk::now();
      b->free_shard(tok.value().remote_addr, size);
      auto f = Clk::now();
      t_res.push_back(ms(a, c));
      t_wr.push_back(ms(c, d));
      t_com.push_back(ms(d, e));
      t_free.push_back(ms(e, f));
    }
    auto avg = [](const std::vector<double>& v) { return v.empty() ? 0.0 : std::accumulate(v.begin(), v.end(), 0.0) / static_cast<double>(v.size()); };
    const size_t done = t_res.size();  // never divide by a zero iteration count (reference bug #15)
    std::printf("%s backend, %zu ops x %llu B: reserve %.4f ms | write %.4f ms | commit %.4f ms | free %.4f ms | %.0f lifecycle ops/s\n",
                args.get("class", "NVME").c_str(), done, static_cast<unsigned long long>(size), avg(t_res), avg(t_wr), avg(t_com), avg(t_free),
                done ? 1000.0 / (avg(t_res) + avg(t_wr) + avg(t_com) + avg(t_free)) : 0.0);
    return 0;
  }
  auto hp = split_host_port(args.get("keystone", "127.0.0.1:9090"));
  if (!hp) return 2;
  client::BlackbirdClientOptions o;
  o.keystone_host = hp->first;
  o.keystone_port = static_cast<uint16_t>(hp->second);
  o.io_parallelism = static_cast<size_t>(args.num("parallelism", 4));
  client::BlackbirdClient cl(o);
  auto t_setup = Clk::now();
  if (cl.connect() != ErrorCode::OK) {
    std::fprintf(stderr, "cannot connect to keystone\n");
    return 1;
  }
  const double setup_ms = ms(t_setup, Clk::now());
  const size_t size = static_cast<size_t>(args.num("size", 1 << 20));
  const int iters = static_cast<int>(args.num("iterations", 10));
  const int batch = static_cast<int>(std::max<long long>(1, args.num("batch", 1)));
  WorkerConfig cfg;
  cfg.replication_factor = static_cast<size_t>(args.num("replicas", 1));
  cfg.max_workers_per_copy = static_cast<size_t>(args.num("max-workers", 1));
  if (args.get("checksum") == "crc32c") cfg.checksum = ChecksumAlgo::CRC32C;
  if (args.get("checksum") == "none") cfg.checksum = ChecksumAlgo::NONE;
  std::vector<uint8_t> data(size);
  std::mt19937_64 rng(1);
  for (size_t i = 0; i + 8 <= size; i += 8) {
    uint64_t v = rng();
    std::memcpy(&data[i], &v, 8);
  }
  const std::string prefix = args.get("key-prefix", "bench-" + std::to_string(Clk::now().time_since_epoch().count()));
  std::vector<double> wr, rd;
  int failures = 0;
  auto t_total = Clk::now();
  for (int it = 0; it < iters; ++it) {
    std::vector<ObjectKey> keys;
    std::vector<const uint8_t*> ptrs;
    std::vector<size_t> sizes;
    for (int b = 0; b < batch; ++b) {
      keys.push_back(prefix + "-" + std::to_string(it) + "-" + std::to_string(b));
      ptrs.push_back(data.data());
      sizes.push_back(size);
    }
    auto a = Clk::now();
    auto ecs = batch == 1 ? std::vector<ErrorCode>{cl.put(keys[0], data.data(), size, cfg)} : cl.batch_put(keys, ptrs, sizes, cfg);
    auto b2 = Clk::now();
    bool ok = std::all_of(ecs.begin(), ecs.end(), [](ErrorCode e) { return e == ErrorCode::OK; });
    if (ok) {
      if (batch == 1) ok = cl.get(keys[0]).ok();
      else {
        auto got = cl.batch_get(keys);
        ok = std::all_of(got.begin(), got.end(), [](const auto& r) { return r.ok(); });
      }
    }
    auto c = Clk::now();
    if (!ok) {
      ++failures;
    } else {
      wr.push_back(ms(a, b2));
      rd.push_back(ms(b2, c));
    }
    cl.batch_remove(keys);
  }
  const double total_ms = ms(t_total, Clk::now());
  auto avg = [](const std::vector<double>& v) { return v.empty() ? 0.0 : std::accumulate(v.begin(), v.end(), 0.0) / static_cast<double>(v.size()); };
  const double bytes = static_cast<double>(size) * batch;
  std::printf("{\"size\": %zu, \"batch\": %d, \"iterations\": %zu, \"failures\": %d, \"setup_ms\": %.3f, \"total_ms\": %.3f, "
              "\"write_avg_ms\": %.4f, \"write_p50_ms\": %.4f, \"write_p99_ms\": %.4f, \"write_MiBps\": %.1f, "
              "\"read_avg_ms\": %.4f, \"read_p50_ms\": %.4f, \"read_p99_ms\": %.4f, \"read_MiBps\": %.1f}\n",
              size, batch, wr.size(), failures, setup_ms, total_ms, avg(wr), pct(wr, 0.5), pct(wr, 0.99),
              wr.empty() ? 0.0 : bytes * 1000.0 / (avg(wr) * 1048576.0), avg(rd), pct(rd, 0.5), pct(rd, 0.99),
              rd.empty() ? 0.0 : bytes * 1000.0 / (avg(rd) * 1048576.0));
  return failures ? 1 : 0;
}
