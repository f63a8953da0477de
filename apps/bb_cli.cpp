// bb-cli: command-line client (reference clients/ucx_client.cpp: put + get + compare with
// timings; examples/simple_client_test.cpp: connectivity + /metrics smoke).
//   bb-cli --keystone 127.0.0.1:9090 put KEY FILE [--replicas R] [--max-workers W] [--ttl-ms T] [--class RAM_CPU]
//   bb-cli get KEY [OUTFILE] | exists KEY | remove KEY | migrate KEY CLASS | where KEY | pools | stats | smoke [--size N] | metrics --http 127.0.0.1:9091
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iterator>
#include <random>

#include "apps/cli_util.h"
#include "net/tcp.h"
#include "common/tenant.h"
#include "client/blackbird_client.h"
#include "common/log.h"

using namespace bb;

namespace {
double ms_since(std::chrono::steady_clock::time_point t0) {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
const char* name(ErrorCode ec) {
  static thread_local std::string s;
  s = std::string(to_string(ec));
  return s.c_str();
}
}  // namespace

int main(int argc, char** argv) {
  auto args = bbapp::parse_args(argc, argv);
  if (args.has("auth-token")) bb::net::set_cluster_token(args.get("auth-token"));  // else BB_AUTH_TOKEN / config
  if (args.has("encrypt-transport")) bb::net::set_transport_encryption(true);  // else BB_ENCRYPT_TRANSPORT / config
  if (args.has("auth-token-ro")) bb::net::set_cluster_token_ro(args.get("auth-token-ro"));  // else BB_AUTH_TOKEN_RO / config
  if (args.has("http-token")) bb::net::set_http_token(args.get("http-token"));  // else BB_HTTP_TOKEN / config: bearer token of /metrics and /stats
  if (args.has("tenant")) bb::set_client_tenant(args.get("tenant"), args.get("tenant-secret"));  // else BB_TENANT / BB_TENANT_SECRET (common/tenant.h)
  if (args.positional.empty() || args.has("help")) {
    std::printf("usage: bb-cli [--keystone host:port] <put KEY FILE | get KEY [OUT] | exists KEY | remove KEY | where KEY | ls [PREFIX] | rm-prefix PREFIX | migrate KEY CLASS | stats | pools | workers | remove-worker ID | drain-worker ID | scrub [PREFIX] [MAX] | compact POOL | tenants | smoke | metrics --http host:port> [--auth-token T | --auth-token-ro T | --tenant NAME --tenant-secret S] [--encrypt-transport]\n");
    return args.has("help") ? 0 : 2;
  }
  const std::string cmd = args.positional[0];
  if (cmd == "metrics") {
    auto hp = split_host_port(args.get("http", "127.0.0.1:9091"));
    if (!hp) return 2;
    int status = 0;
    auto r = net::http_get(hp->first, static_cast<uint16_t>(hp->second), args.get("path", "/metrics"), &status);
    if (!r.ok()) {
      std::fprintf(stderr, "http error: %s\n", name(r.error()));
      return 1;
    }
    std::printf("%s", r.value().c_str());
    return status == 200 ? 0 : 1;
  }
  // --keystone takes one endpoint or, for an HA pair, a comma-separated list; the client follows the leader.
  client::BlackbirdClientOptions o;
  {
    const std::string list = args.get("keystone", "127.0.0.1:9090");
    size_t pos = 0;
    while (pos <= list.size()) {
      const size_t comma = std::min(list.find(',', pos), list.size());
      if (comma > pos) {
        if (!split_host_port(list.substr(pos, comma - pos))) return 2;
        o.keystone_endpoints.push_back(list.substr(pos, comma - pos));
      }
      pos = comma + 1;
    }
    if (o.keystone_endpoints.empty()) return 2;
  }
  o.node_id = args.get("node-id");
  o.io_parallelism = static_cast<size_t>(args.num("parallelism", 4));
  client::BlackbirdClient cl(o);
  auto t0 = std::chrono::steady_clock::now();
  ErrorCode ec = cl.connect();
  if (ec != ErrorCode::OK) {
    std::fprintf(stderr, "cannot connect to keystone %s: %s\n", args.get("keystone", "127.0.0.1:9090").c_str(), name(ec));
    return 1;
  }
  const double connect_ms = ms_since(t0);
  WorkerConfig cfg;
  cfg.replication_factor = static_cast<size_t>(args.num("replicas", 1));
  cfg.max_workers_per_copy = static_cast<size_t>(args.num("max-workers", 1));
  cfg.ttl_ms = static_cast<uint64_t>(args.num("ttl-ms", 30 * 60 * 1000));
  cfg.enable_soft_pin = args.has("soft-pin");
  if (args.has("class"))
    if (auto sc = parse_storage_class(args.get("class"))) cfg.preferred_classes = {*sc};
  if (args.get("checksum") == "crc32c") cfg.checksum = ChecksumAlgo::CRC32C;
  if (args.get("checksum") == "xxh3") cfg.checksum = ChecksumAlgo::XXH3;

  if (cmd == "put" && args.positional.size() >= 3) {
    std::ifstream f(args.positional[2], std::ios::binary);
    if (!f) {
      std::fprintf(stderr, "cannot read %s\n", args.positional[2].c_str());
      return 2;
    }
    std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    t0 = std::chrono::steady_clock::now();
    ec = cl.put(args.positional[1], data, cfg);
    const double ms = ms_since(t0);
    std::printf("put %s: %s (%zu bytes, %.3f ms, %.1f MB/s)\n", args.positional[1].c_str(), name(ec), data.size(), ms, data.size() / ms / 1e3);
    return ec == ErrorCode::OK ? 0 : 1;
  }
  if (cmd == "get" && args.positional.size() >= 2) {
    t0 = std::chrono::steady_clock::now();
    auto r = cl.get(args.positional[1]);
    const double ms = ms_since(t0);
    if (!r.ok()) {
      std::fprintf(stderr, "get %s: %s\n", args.positional[1].c_str(), name(r.error()));
      return 1;
    }
    if (args.positional.size() >= 3) {
      std::ofstream o2(args.positional[2], std::ios::binary);
      o2.write(reinterpret_cast<const char*>(r.value().data()), static_cast<std::streamsize>(r.value().size()));
    }
    std::printf("get %s: OK (%zu bytes, %.3f ms, %.1f MB/s, checksum verified)\n", args.positional[1].c_str(), r.value().size(), ms, r.value().size() / ms / 1e3);
    return 0;
  }
  if (cmd == "exists" && args.positional.size() >= 2) {
    auto r = cl.object_exists(args.positional[1]);
    std::printf("%s\n", r.ok() && r.value() ? "true" : "false");
    return r.ok() && r.value() ? 0 : 1;
  }
  if (cmd == "remove" && args.positional.size() >= 2) {
    ec = cl.remove(args.positional[1]);
    std::printf("remove %s: %s\n", args.positional[1].c_str(), name(ec));
    return ec == ErrorCode::OK ? 0 : 1;
  }
  if (cmd == "migrate" && args.positional.size() >= 3) {
    auto sc = parse_storage_class(args.positional[2]);
    if (!sc) {
      std::fprintf(stderr, "unknown storage class %s\n", args.positional[2].c_str());
      return 2;
    }
    ec = cl.migrate(args.positional[1], *sc);
    std::printf("migrate %s -> %s: %s\n", args.positional[1].c_str(), args.positional[2].c_str(), name(ec));
    return ec == ErrorCode::OK ? 0 : 1;
  }
  if (cmd == "pools") {  // admin introspection: every registered pool with live usage
    auto pools = cl.keystone().get_memory_pools();
    if (!pools.ok()) return 1;
    Json arr = Json::array();
    for (const auto& p : pools.value()) arr.push_back(to_json(p));
    std::printf("%s\n", arr.dump(2).c_str());
    return 0;
  }
  if (cmd == "ls") {  // bb-cli ls [PREFIX] [--limit N] [--after KEY]
    auto v = cl.keystone().list_objects(args.positional.size() >= 2 ? args.positional[1] : "", static_cast<size_t>(args.num("limit", 1000)),
                                        args.get("after"));
    if (!v.ok()) return 1;
    for (const auto& o : v.value())
      std::printf("%12llu  x%u  %-10s %s\n", static_cast<unsigned long long>(o.size), o.copies, std::string(to_string(o.tier)).c_str(), o.key.c_str());
    return 0;
  }
  if (cmd == "rm-prefix" && args.positional.size() >= 2) {  // remove every object under a prefix, page by page
    size_t removed = 0;
    while (true) {
      auto v = cl.keystone().list_objects(args.positional[1], 1000, "");
      if (!v.ok()) return 1;
      if (v.value().empty()) break;
      std::vector<ObjectKey> keys;
      for (const auto& o : v.value()) keys.push_back(o.key);
      for (ErrorCode e : cl.batch_remove(keys)) removed += e == ErrorCode::OK ? 1 : 0;
      if (v.value().size() < 1000) break;
    }
    std::printf("removed %zu objects under %s\n", removed, args.positional[1].c_str());
    return 0;
  }
  if (cmd == "compact" && args.positional.size() >= 2) {  // defragment a pool: bb-cli compact POOL [--max-moves N]
    auto r = cl.keystone().compact_pool(args.positional[1], static_cast<size_t>(args.num("max-moves", 64)));
    if (!r.ok()) {
      std::printf("compact %s: %s\n", args.positional[1].c_str(), name(r.error()));
      return 1;
    }
    std::printf("compact %s: moved %zu objects\n", args.positional[1].c_str(), r.value());
    return 0;
  }
  if (cmd == "workers") {  // admin introspection: registered workers, heartbeat age, their pools
    auto ws = cl.keystone().get_workers_info();
    if (!ws.ok()) return 1;
    Json arr = Json::array();
    for (const auto& w : ws.value()) {
      Json j = Json::object();
      j["worker_id"] = w.worker_id;
      j["node_id"] = w.node_id;
      j["endpoint"] = w.endpoint;
      j["heartbeat_age_ms"] = static_cast<int64_t>(w.heartbeat_age_ms);
      Json ps = Json::array();
      for (const auto& p : w.pools) ps.push_back(p);
      j["pools"] = ps;
      arr.push_back(j);
    }
    std::printf("%s\n", arr.dump(2).c_str());
    return 0;
  }
  if (cmd == "remove-worker" && args.positional.size() >= 2) {  // decommission: copies there are invalidated and re-replicated
    ec = cl.keystone().remove_worker(args.positional[1]);
    std::printf("remove-worker %s: %s\n", args.positional[1].c_str(), name(ec));
    return ec == ErrorCode::OK ? 0 : 1;
  }
  if (cmd == "drain-worker" && args.positional.size() >= 2) {  // graceful decommission: move everything off first, then remove
    auto r = cl.keystone().drain_worker(args.positional[1]);
    if (!r.ok()) {
      std::printf("drain-worker %s: %s\n", args.positional[1].c_str(), name(r.error()));
      return 1;
    }
    std::printf("drain-worker %s: OK, %zu objects moved\n", args.positional[1].c_str(), r.value());
    return 0;
  }
  if (cmd == "tenants") {  // what each tenant holds against its budget (a tenant sees its own line)
    auto r = cl.keystone().tenant_usage();
    if (!r.ok()) {
      std::printf("tenants: %s\n", name(r.error()));
      return 1;
    }
    std::printf("%-24s %16s %10s %16s %12s\n", "tenant", "used_bytes", "objects", "quota_bytes", "max_objects");
    for (const auto& u : r.value())
      std::printf("%-24s %16llu %10llu %16llu %12llu\n", u.name.c_str(), static_cast<unsigned long long>(u.used_bytes), static_cast<unsigned long long>(u.objects),
                  static_cast<unsigned long long>(u.quota_bytes), static_cast<unsigned long long>(u.max_objects));
    return 0;
  }
  if (cmd == "scrub") {  // re-hash stored copies where they lie; replace the ones that no longer match their digest
    const std::string prefix = args.positional.size() >= 2 ? args.positional[1] : "";
    const size_t max_objects = args.positional.size() >= 3 ? std::strtoull(args.positional[2].c_str(), nullptr, 10) : 0;
    auto r = cl.keystone().scrub(prefix, max_objects);
    if (!r.ok()) {
      std::printf("scrub: %s\n", name(r.error()));
      return 1;
    }
    const auto& v = r.value();
    std::printf("scrub: %llu objects, %llu copies hashed, %llu corrupt, %llu healed, %llu unrecoverable, %llu unreachable\n",
                (unsigned long long)v.objects, (unsigned long long)v.copies, (unsigned long long)v.corrupt, (unsigned long long)v.healed,
                (unsigned long long)v.unrecoverable, (unsigned long long)v.unreachable);
    return v.unrecoverable ? 2 : 0;
  }
  if (cmd == "where" && args.positional.size() >= 2) {  // placement of an object: copy -> shards (pool, worker, tier, digest)
    auto copies = cl.get_workers(args.positional[1]);
    if (!copies.ok()) {
      std::printf("where %s: %s\n", args.positional[1].c_str(), name(copies.error()));
      return 1;
    }
    for (const auto& c : copies.value())
      for (size_t i = 0; i < c.shards.size(); ++i) {
        const auto& s = c.shards[i];
        std::printf("copy %u shard %zu: pool=%s worker=%s tier=%s bytes=%llu %s=%016llx\n", c.copy_index, i, s.pool_id.c_str(), s.worker_id.c_str(),
                    std::string(to_string(s.storage_class)).c_str(), static_cast<unsigned long long>(s.length),
                    std::string(to_string(s.checksum_algo)).c_str(), static_cast<unsigned long long>(s.checksum));
      }
    return 0;
  }
  if (cmd == "stats") {
    auto st = cl.cluster_stats();
    if (!st.ok()) return 1;
    std::printf("%s\n", to_json(st.value()).dump(2).c_str());
    return 0;
  }
  if (cmd == "smoke") {
    // put -> get -> compare -> placements -> remove, with timings (reference clients/ucx_client.cpp:188-334)
    const size_t n = static_cast<size_t>(args.num("size", 64));
    std::vector<uint8_t> data(n);
    std::mt19937 rng(42);
    for (auto& b : data) b = static_cast<uint8_t>(rng());
    const std::string key = args.get("key", "smoke-" + std::to_string(std::chrono::steady_clock::now().time_since_epoch().count()));
    t0 = std::chrono::steady_clock::now();
    ec = cl.put(key, data, cfg);
    const double put_ms = ms_since(t0);
    if (ec != ErrorCode::OK) {
      std::fprintf(stderr, "smoke put failed: %s\n", name(ec));
      return 1;
    }
    t0 = std::chrono::steady_clock::now();
    auto back = cl.get(key);
    const double get_ms = ms_since(t0);
    const bool same = back.ok() && back.value() == data;
    auto placed = cl.get_workers(key);
    size_t shards = 0;
    if (placed.ok())
      for (const auto& c : placed.value()) shards += c.shards.size();
    cl.remove(key);
    std::printf("connect %.3f ms | put %.3f ms (%.2f MB/s) | get %.3f ms (%.2f MB/s) | copies %zu shards %zu | verify %s\n", connect_ms, put_ms,
                n / put_ms / 1e3, get_ms, n / get_ms / 1e3, placed.ok() ? placed.value().size() : 0, shards, same ? "PASS" : "FAIL");
    return same ? 0 : 1;
  }
  std::fprintf(stderr, "unknown command\n");
  return 2;
}
