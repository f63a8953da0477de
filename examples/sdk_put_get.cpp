// C++ SDK example: connect to a running Keystone (scripts/start_cluster.sh), put an object with two replicas,
// read it back (every shard digest-checked, replica fail-over), use the batch API, print cluster stats.
//   bin/bb-example-sdk-put-get [host:port]
#include <cstdio>
#include <numeric>
#include <string>
#include <vector>

#include "client/blackbird_client.h"

using namespace bb;

int main(int argc, char** argv) {
  const std::string addr = argc > 1 ? argv[1] : "127.0.0.1:9090";
  auto hp = split_host_port(addr);
  if (!hp) {
    std::fprintf(stderr, "usage: %s host:port\n", argv[0]);
    return 2;
  }
  client::BlackbirdClientOptions opts;
  opts.keystone_host = hp->first;
  opts.keystone_port = static_cast<uint16_t>(hp->second);
  opts.node_id = "example-client";
  client::BlackbirdClient cl(opts);
  if (cl.connect() != ErrorCode::OK) {
    std::fprintf(stderr, "cannot reach the keystone at %s\n", addr.c_str());
    return 1;
  }
  WorkerConfig cfg;
  cfg.replication_factor = 2;
  cfg.max_workers_per_copy = 1;
  cfg.ttl_ms = 60'000;
  cfg.checksum = ChecksumAlgo::CRC32C;
  std::vector<uint8_t> blob(1 << 20);
  std::iota(blob.begin(), blob.end(), uint8_t{0});
  ErrorCode ec = cl.put("example/blob", blob, cfg);
  if (ec == ErrorCode::INSUFFICIENT_SPACE) {  // single-worker cluster: one copy is all it can hold
    cfg.replication_factor = 1;
    ec = cl.put("example/blob", blob, cfg);
  }
  if (ec != ErrorCode::OK) {
    std::fprintf(stderr, "put failed: %s\n", std::string(to_string(ec)).c_str());
    return 1;
  }
  auto back = cl.get("example/blob");
  if (!back.ok() || back.value() != blob) {
    std::fprintf(stderr, "get failed or data mismatch\n");
    return 1;
  }
  auto copies = cl.get_workers("example/blob");
  std::printf("example/blob: %zu bytes, %zu copies, first shard on %s (crc32c %08llx)\n", blob.size(), copies.value().size(),
              copies.value()[0].shards[0].worker_id.c_str(), static_cast<unsigned long long>(copies.value()[0].shards[0].checksum));
  std::vector<ObjectKey> keys;
  std::vector<std::vector<uint8_t>> values;
  for (int i = 0; i < 4; ++i) {
    keys.push_back("example/batch/" + std::to_string(i));
    values.emplace_back(4096, static_cast<uint8_t>(i));
  }
  std::vector<const uint8_t*> ptrs;
  std::vector<size_t> sizes;
  for (const auto& v : values) ptrs.push_back(v.data()), sizes.push_back(v.size());
  for (ErrorCode e : cl.batch_put(keys, ptrs, sizes, cfg))
    if (e != ErrorCode::OK) return 1;
  auto got = cl.batch_get(keys);
  for (size_t i = 0; i < keys.size(); ++i)
    if (!got[i].ok() || got[i].value() != values[i]) return 1;
  auto st = cl.cluster_stats();
  if (st.ok())
    std::printf("cluster: %zu workers, %zu pools, %zu objects, %.1f%% used\n", st.value().total_workers, st.value().total_memory_pools,
                st.value().total_objects, 100.0 * st.value().avg_utilization);
  keys.push_back("example/blob");
  cl.batch_remove(keys);
  std::printf("sdk example OK\n");
  return 0;
}
