// WorkerService: owns storage pools, serves their bytes, advertises them.
//
// Parity: reference include/blackbird/worker/worker_service.h:21-155 (WorkerServiceConfig,
// load_worker_config_from_file, initialize/start/stop, add_storage_pool,
// create_storage_pools_from_config, get_stats) and the registration protocol of
// worker_service.cpp:399-516 (worker JSON, per-pool JSON, TTL heartbeat key).
//
// Differences by design: the worker actually serves requests — a framed-RPC data server
// exposes WRITE / READ / CHECKSUM on every pool, which is the data path for host tiers and the
// slow path for GPU slabs (the fast path is the fused NVLink kernel, fabric/).  Heartbeat
// interval and TTL come from the config (reference hard-codes 5 s / 10 s, bug #12) and reuse
// one lease.  Without a coordination daemon a worker registers directly with a keystone.
#pragma once
#include <atomic>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common/cxl_config.h"
#include "coord/coord.h"
#include "net/tcp.h"
#include "rpc/rpc_service.h"
#include "worker/storage_backend.h"

namespace bb::worker {

struct StoragePoolConfig {
  std::string pool_id;
  StorageClass storage_class = StorageClass::RAM_CPU;
  uint64_t size_bytes = 0;
  std::string mount_path;   // disk tiers; dax device for CXL
  int gpu_device_id = 0;
  int numa_node = -1;
  uint32_t queue_depth = 64;
  bool pin_memory = false;  // DRAM tier: register with CUDA so fused kernels can move data to / from it
  bool shared_memory = false;  // DRAM tier: memfd-backed, mappable by GPU clients of other processes on this host
  bool encrypt_at_rest = false;  // file-backed tiers: AES-256-CTR of the pool bytes, keyed from WorkerServiceConfig::at_rest_key
  CxlMemoryPoolConfig cxl;  // CXL tiers: per-pool `config:` block
};

struct WorkerServiceConfig {
  std::string worker_id;
  std::string node_id;
  std::string cluster_id = DEFAULT_CLUSTER_ID;
  std::string etcd_endpoints;       // coordination endpoints ("" = none)
  std::string keystone_address;     // direct registration when no coordination store is used
  std::string rpc_endpoint = "0.0.0.0:0";
  std::string ucx_endpoint = "127.0.0.1:0";  // data server listen address (reference name kept)
  std::vector<std::string> interconnects{"tcp"};
  double max_bw_gbps = 0.0;
  int numa_node = -1;
  std::string version = "1.0.0";
  int64_t lease_ttl_sec = 10;
  int64_t heartbeat_interval_sec = 5;
  // How often expired shard reservations are reclaimed and reported to the Keystone (reference worker_service.h:37,
  // where the knob exists but nothing reads it).
  int64_t allocation_poll_interval_ms = 1000;
  std::string fabric_domain;        // e.g. "nvswitch-0"
  // HTTP port of the worker's own observability endpoint (/metrics, /healthz, /stats); -1 = disabled, 0 = ephemeral.
  int http_metrics_port = -1;
  std::string auth_token;  // shared cluster token (net/tcp.h); empty = BB_AUTH_TOKEN / open cluster
  bool encrypt_transport = false;  // secure mode of the RPC protocol (net/tcp.h)
  std::string auth_token_ro;       // read-only members' token (net/tcp.h)
  std::string http_auth_token;     // bearer token of the worker's /metrics and /stats (net/tcp.h); BB_HTTP_TOKEN
  std::string audit_log;           // audit trail of this worker's data server (refused handshakes, denied methods); BB_AUDIT_LOG
  std::string tenants_file;        // tenant table (common/tenant.h): the data server admits tenants for reads and writes; BB_TENANTS_FILE
  std::string at_rest_key;         // passphrase of pools with encrypt_at_rest (BB_AT_REST_KEY is the default)
  CxlTransportConfig transport;     // `transport:` block (cxl_worker.yaml); drives the advertised interconnects
  bool has_transport = false;
  std::vector<TierRule> preferred_tiers;  // `allocation.preferred_tiers` (forwarded to the keystone as a hint)
  std::vector<StoragePoolConfig> storage_pools;
};

Result<WorkerServiceConfig> worker_config_from_json(const Json& root, std::string* err = nullptr);
// Throws std::runtime_error on unreadable / invalid files (as the reference does).
WorkerServiceConfig load_worker_config_from_file(const std::string& path);

// D_PULL: the destination worker copies a shard out of a *peer worker's* GPU slab itself (opens the peer's CUDA IPC
// handle, one fused-kernel launch over NVLink) -- the re-replication / repair path between GPU-tier workers.
// D_RESERVE / D_COMMIT / D_ABORT / D_FREE: the reservation protocol between the Keystone and the workers
// (StorageBackend::reserve_shard_at -> commit | abort, later free), reference storage_backend.h:46-126.
enum DataMethod : uint32_t { D_WRITE = 1, D_READ = 2, D_CHECKSUM = 3, D_STATS = 4, D_COPY = 5, D_PULL = 6, D_RESERVE = 7, D_COMMIT = 8,
                             D_ABORT = 9, D_FREE = 10 };

class WorkerService {
 public:
  explicit WorkerService(const WorkerServiceConfig& config, std::shared_ptr<coord::CoordService> coord = nullptr,
                         std::shared_ptr<rpc::KeystoneApi> keystone = nullptr);
  ~WorkerService();
  WorkerService(const WorkerService&) = delete;
  WorkerService& operator=(const WorkerService&) = delete;

  ErrorCode add_storage_pool(const std::string& pool_id, std::unique_ptr<StorageBackend> backend);
  ErrorCode create_storage_pools_from_config();
  ErrorCode initialize();
  ErrorCode start();
  void stop();
  bool is_running() const { return running_.load(); }

  Json get_stats() const;
  // Prometheus text exposition of the per-pool counters (capacity / used / reservations / bytes moved / I/O errors /
  // fused tier moves) and of the data server; served on /metrics when `http_metrics_port` >= 0.
  std::string metrics_text() const;
  uint16_t http_port() const { return http_server_.port(); }
  StorageBackend* backend(const std::string& pool_id);
  std::vector<MemoryPool> advertised_pools() const;
  uint16_t data_port() const { return data_server_.port(); }
  std::string data_endpoint() const;
  const WorkerServiceConfig& config() const { return config_; }
  // Extra registration attributes for a pool (e.g. the CUDA IPC handle of a GPU slab).
  void set_pool_rkey_hex(const std::string& pool_id, const std::string& hex);

  // Runs one reservation sweep now (tests); returns the number of expired reservations reclaimed.
  size_t reap_reservations() { return reap_reservations_once(); }
  // Fault injection: "drop_heartbeat" stops refreshing the lease, "" clears.
  void inject_fault(const std::string& fault);

 private:
  void register_data_handlers();
  ErrorCode register_all();
  void heartbeat_loop();
  void reservation_reaper_loop();
  size_t reap_reservations_once();
  MemoryPool describe_pool(const std::string& pool_id, const StorageBackend& b) const;
  std::string cluster_prefix() const { return "/blackbird/clusters/" + config_.cluster_id + "/"; }

  WorkerServiceConfig config_;
  std::shared_ptr<coord::CoordService> coord_;
  std::shared_ptr<rpc::KeystoneApi> keystone_;
  mutable std::mutex pools_mu_;
  std::map<std::string, std::unique_ptr<StorageBackend>> pools_;
  std::map<std::string, StoragePoolConfig> pool_cfg_;
  std::map<std::string, std::string> pool_rkey_hex_;
  net::RpcServer data_server_;
  net::HttpServer http_server_;
  std::atomic<bool> running_{false};
  std::atomic<bool> initialized_{false};
  std::atomic<bool> drop_heartbeat_{false};
  std::thread heartbeat_thread_;
  std::thread reaper_thread_;
  std::atomic<uint64_t> reservations_expired_{0};
  std::mutex sleep_mu_;
  std::condition_variable sleep_cv_;
  std::atomic<uint64_t> heartbeats_sent_{0};
};

}  // namespace bb::worker
