// RPC + HTTP façade of the Keystone and its client stub.
//
// Parity: reference include/blackbird/rpc/rpc_service.h:28-274 — 14 rpc_* handlers registered at
// rpc_service.cpp:369-382 (object_exists, get_workers, put_start, put_complete, put_cancel,
// remove_object, remove_all_objects, get_cluster_stats, get_view_version and the five batch_*),
// typed in-process wrappers (:69-167), metrics HTTP server (:212-226, which has no routes in
// the reference — /metrics, /healthz and /stats are real here), create_and_start_keystone (:274).
// Extra methods: batch_remove_object, client_register / client_ping (sessions), pool and worker
// introspection / registration for deployments without a coordination daemon.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "keystone/keystone_service.h"
#include "net/tcp.h"

namespace bb::rpc {

enum Method : uint32_t {
  M_OBJECT_EXISTS = 1, M_GET_WORKERS, M_PUT_START, M_PUT_COMPLETE, M_PUT_CANCEL, M_REMOVE_OBJECT,
  M_REMOVE_ALL_OBJECTS, M_GET_CLUSTER_STATS, M_GET_VIEW_VERSION, M_BATCH_OBJECT_EXISTS,
  M_BATCH_GET_WORKERS, M_BATCH_PUT_START, M_BATCH_PUT_COMPLETE, M_BATCH_PUT_CANCEL,
  M_BATCH_REMOVE_OBJECT, M_CLIENT_REGISTER, M_CLIENT_PING, M_GET_MEMORY_POOLS,
  M_REGISTER_WORKER, M_REGISTER_MEMORY_POOL, M_WORKER_HEARTBEAT, M_REMOVE_WORKER,
  M_MIGRATE_OBJECT, M_GET_WORKERS_INFO, M_LIST_OBJECTS, M_COMPACT_POOL, M_DRAIN_WORKER, M_SCRUB, M_TENANT_USAGE,
};

// The keystone surface a client needs; implemented in-process and over TCP.
class KeystoneApi {
 public:
  virtual ~KeystoneApi() = default;
  virtual Result<bool> object_exists(const ObjectKey& key) = 0;
  virtual Result<std::vector<CopyPlacement>> get_workers(const ObjectKey& key) = 0;
  virtual Result<std::vector<CopyPlacement>> put_start(const ObjectKey& key, size_t size, const WorkerConfig& cfg) = 0;
  virtual ErrorCode put_complete(const ObjectKey& key, const keystone::ShardChecksums& sums) = 0;
  virtual ErrorCode put_cancel(const ObjectKey& key) = 0;
  virtual ErrorCode remove_object(const ObjectKey& key) = 0;
  virtual Result<size_t> remove_all_objects() = 0;
  virtual Result<ClusterStats> get_cluster_stats() = 0;
  virtual Result<ViewVersionId> get_view_version() = 0;
  virtual std::vector<Result<bool>> batch_object_exists(const std::vector<ObjectKey>& keys) = 0;
  virtual std::vector<Result<std::vector<CopyPlacement>>> batch_get_workers(const std::vector<ObjectKey>& keys) = 0;
  virtual std::vector<Result<std::vector<CopyPlacement>>> batch_put_start(const std::vector<keystone::PutStartItem>& items) = 0;
  virtual std::vector<ErrorCode> batch_put_complete(const std::vector<ObjectKey>& keys,
                                                    const std::vector<keystone::ShardChecksums>& sums) = 0;
  virtual std::vector<ErrorCode> batch_put_cancel(const std::vector<ObjectKey>& keys) = 0;
  virtual std::vector<ErrorCode> batch_remove_object(const std::vector<ObjectKey>& keys) = 0;
  virtual Result<std::vector<MemoryPool>> get_memory_pools() = 0;
  virtual Result<std::string> client_register(const std::string& node_id) = 0;
  virtual Result<ViewVersionId> client_ping(const std::string& client_id) = 0;
  virtual ErrorCode register_worker(const WorkerRecord& rec) = 0;
  virtual ErrorCode register_memory_pool(const MemoryPool& pool) = 0;
  virtual ErrorCode worker_heartbeat(const WorkerId& id) = 0;
  // Explicit promotion / demotion of an object to another tier (all copies).
  virtual ErrorCode migrate_object(const ObjectKey& key, StorageClass target) = 0;
  // Admin introspection / maintenance (reference keystone_service.h:105-165: get_workers_info, remove_worker).
  struct WorkerSummary {
    WorkerId worker_id;
    NodeId node_id;
    std::string endpoint;
    int64_t heartbeat_age_ms = 0;
    std::vector<MemoryPoolId> pools;
  };
  virtual Result<std::vector<WorkerSummary>> get_workers_info() = 0;
  // Listing (extension): complete objects under `prefix`, key order, paginated with `start_after`.
  virtual Result<std::vector<keystone::KeystoneService::ListedObject>> list_objects(const std::string& prefix, size_t limit,
                                                                                    const std::string& start_after) = 0;
  virtual ErrorCode remove_worker(const WorkerId& id) = 0;
  // Defragments a pool by re-placing its highest objects into lower holes; returns the number of objects moved.
  virtual Result<size_t> compact_pool(const MemoryPoolId& pool, size_t max_moves) = 0;
  virtual Result<size_t> drain_worker(const WorkerId& id) = 0;  // move everything off a worker, then remove it
  // re-hash stored copies at their workers and replace the ones that rotted (KeystoneService::scrub)
  virtual Result<keystone::ScrubReport> scrub(const std::string& prefix, size_t max_objects) = 0;
  // what each tenant holds against its budget (a tenant sees only itself; common/tenant.h)
  virtual Result<std::vector<keystone::TenantUsage>> tenant_usage() = 0;
  // identity used for locality-aware placement and session ownership
  void set_identity(std::string client_id, std::string node_id) {
    client_id_ = std::move(client_id);
    node_id_ = std::move(node_id);
  }
  const std::string& client_id() const { return client_id_; }
  const std::string& node_id() const { return node_id_; }

 protected:
  std::string client_id_, node_id_;
};

class LocalKeystoneApi : public KeystoneApi {
 public:
  explicit LocalKeystoneApi(std::shared_ptr<keystone::KeystoneService> ks) : ks_(std::move(ks)) {}
  Result<bool> object_exists(const ObjectKey& key) override { return ks_->object_exists(key); }
  Result<std::vector<CopyPlacement>> get_workers(const ObjectKey& key) override { return ks_->get_workers(key); }
  Result<std::vector<CopyPlacement>> put_start(const ObjectKey& key, size_t size, const WorkerConfig& cfg) override {
    return ks_->put_start(key, size, cfg, client_id_, node_id_);
  }
  ErrorCode put_complete(const ObjectKey& key, const keystone::ShardChecksums& sums) override { return ks_->put_complete(key, sums); }
  ErrorCode put_cancel(const ObjectKey& key) override { return ks_->put_cancel(key); }
  ErrorCode remove_object(const ObjectKey& key) override { return ks_->remove_object(key); }
  ErrorCode migrate_object(const ObjectKey& key, StorageClass target) override { return ks_->migrate_object(key, target); }
  Result<std::vector<WorkerSummary>> get_workers_info() override {
    std::vector<keystone::WorkerInfo> v;
    ks_->get_workers_info(v);
    std::vector<WorkerSummary> out;
    const auto now = Clock::now();
    for (const auto& w : v)
      out.push_back({w.worker_id, w.node_id, w.endpoint, std::chrono::duration_cast<std::chrono::milliseconds>(now - w.last_heartbeat).count(), w.pools});
    return out;
  }
  ErrorCode remove_worker(const WorkerId& id) override { return ks_->remove_worker(id); }
  Result<size_t> compact_pool(const MemoryPoolId& pool, size_t max_moves) override { return ks_->compact_pool(pool, max_moves); }
  Result<size_t> drain_worker(const WorkerId& id) override { return ks_->drain_worker(id); }
  Result<keystone::ScrubReport> scrub(const std::string& prefix, size_t max_objects) override { return ks_->scrub(prefix, max_objects); }
  Result<std::vector<keystone::TenantUsage>> tenant_usage() override { return ks_->tenant_usage(); }
  Result<std::vector<keystone::KeystoneService::ListedObject>> list_objects(const std::string& prefix, size_t limit,
                                                                            const std::string& start_after) override {
    return ks_->list_objects(prefix, limit, start_after);
  }
  Result<size_t> remove_all_objects() override { return ks_->remove_all_objects(); }
  Result<ClusterStats> get_cluster_stats() override { return ks_->get_cluster_stats(); }
  Result<ViewVersionId> get_view_version() override { return ks_->get_view_version(); }
  std::vector<Result<bool>> batch_object_exists(const std::vector<ObjectKey>& keys) override { return ks_->batch_object_exists(keys); }
  std::vector<Result<std::vector<CopyPlacement>>> batch_get_workers(const std::vector<ObjectKey>& keys) override {
    return ks_->batch_get_workers(keys);
  }
  std::vector<Result<std::vector<CopyPlacement>>> batch_put_start(const std::vector<keystone::PutStartItem>& items) override {
    return ks_->batch_put_start(items, client_id_, node_id_);
  }
  std::vector<ErrorCode> batch_put_complete(const std::vector<ObjectKey>& keys, const std::vector<keystone::ShardChecksums>& sums) override {
    return ks_->batch_put_complete(keys, sums);
  }
  std::vector<ErrorCode> batch_put_cancel(const std::vector<ObjectKey>& keys) override { return ks_->batch_put_cancel(keys); }
  std::vector<ErrorCode> batch_remove_object(const std::vector<ObjectKey>& keys) override { return ks_->batch_remove_object(keys); }
  Result<std::vector<MemoryPool>> get_memory_pools() override {
    std::vector<MemoryPool> v;
    ks_->get_memory_pools(v);
    return v;
  }
  Result<std::string> client_register(const std::string& node_id) override { return ks_->client_register(node_id); }
  Result<ViewVersionId> client_ping(const std::string& id) override { return ks_->client_ping(id); }
  ErrorCode register_worker(const WorkerRecord& rec) override { return ks_->register_worker(rec); }
  ErrorCode register_memory_pool(const MemoryPool& pool) override { return ks_->register_memory_pool(pool); }
  ErrorCode worker_heartbeat(const WorkerId& id) override { return ks_->worker_heartbeat(id); }
  keystone::KeystoneService& service() { return *ks_; }

 private:
  std::shared_ptr<keystone::KeystoneService> ks_;
};

class RpcService {
 public:
  RpcService(std::shared_ptr<keystone::KeystoneService> keystone, const KeystoneConfig& config);
  ~RpcService();
  ErrorCode start();
  void stop();
  bool is_running() const noexcept { return running_; }
  uint16_t rpc_port() const { return rpc_.port(); }
  uint16_t http_port() const { return http_.port(); }
  uint64_t requests_served() const { return rpc_.requests_served(); }
  uint64_t shm_requests_served() const { return rpc_.shm_requests_served(); }  // of which over same-host shared-memory channels
  size_t shm_channels() const { return rpc_.shm_channels(); }
  std::shared_ptr<keystone::KeystoneService> keystone() { return keystone_; }

 private:
  void register_handlers();
  // Object-scoped methods answer NOT_LEADER on a standby (its object map is only recovered when it is elected), which is
  // what tells a multi-endpoint KeystoneRpcClient to move on to the next keystone.
  void leader_only(uint32_t method, net::RpcServer::Handler h);
  std::shared_ptr<keystone::KeystoneService> keystone_;
  KeystoneConfig config_;
  net::RpcServer rpc_;
  net::HttpServer http_;
  bool running_ = false;
};

class KeystoneRpcClient : public KeystoneApi {
 public:
  KeystoneRpcClient() = default;
  ErrorCode connect(const std::string& host, uint16_t port, int timeout_ms = 3000);
  // `host_port` may be a comma-separated list of keystones ("a:9090,b:9090"); see connect_any().
  ErrorCode connect(const std::string& host_port, int timeout_ms = 3000);
  // HA: remembers every endpoint and connects to the first reachable one.  Afterwards a call that fails on the
  // transport (keystone died) or is answered NOT_LEADER (keystone is a standby) moves to the next endpoint and is
  // retried, for up to failover_budget_ms; with a single endpoint behaviour is unchanged (the error is returned).
  // A retried mutation is at-least-once: a put_start whose reply was lost with the old leader can come back
  // OBJECT_ALREADY_EXISTS from the new one.
  ErrorCode connect_any(const std::vector<std::string>& endpoints, int timeout_ms = 3000);
  void set_timeout_ms(int ms) { timeout_ms_ = ms; }
  void set_failover_budget_ms(int ms) { failover_budget_ms_ = ms; }
  bool connected() const;
  std::string active_endpoint() const;
  uint64_t failovers() const { return failovers_.load(); }
  Result<bool> object_exists(const ObjectKey& key) override;
  Result<std::vector<CopyPlacement>> get_workers(const ObjectKey& key) override;
  Result<std::vector<CopyPlacement>> put_start(const ObjectKey& key, size_t size, const WorkerConfig& cfg) override;
  ErrorCode put_complete(const ObjectKey& key, const keystone::ShardChecksums& sums) override;
  ErrorCode put_cancel(const ObjectKey& key) override;
  ErrorCode remove_object(const ObjectKey& key) override;
  ErrorCode migrate_object(const ObjectKey& key, StorageClass target) override;
  Result<std::vector<WorkerSummary>> get_workers_info() override;
  ErrorCode remove_worker(const WorkerId& id) override;
  Result<size_t> compact_pool(const MemoryPoolId& pool, size_t max_moves) override;
  Result<size_t> drain_worker(const WorkerId& id) override;
  Result<keystone::ScrubReport> scrub(const std::string& prefix, size_t max_objects) override;
  Result<std::vector<keystone::TenantUsage>> tenant_usage() override;
  Result<std::vector<keystone::KeystoneService::ListedObject>> list_objects(const std::string& prefix, size_t limit,
                                                                            const std::string& start_after) override;
  Result<size_t> remove_all_objects() override;
  Result<ClusterStats> get_cluster_stats() override;
  Result<ViewVersionId> get_view_version() override;
  std::vector<Result<bool>> batch_object_exists(const std::vector<ObjectKey>& keys) override;
  std::vector<Result<std::vector<CopyPlacement>>> batch_get_workers(const std::vector<ObjectKey>& keys) override;
  std::vector<Result<std::vector<CopyPlacement>>> batch_put_start(const std::vector<keystone::PutStartItem>& items) override;
  std::vector<ErrorCode> batch_put_complete(const std::vector<ObjectKey>& keys, const std::vector<keystone::ShardChecksums>& sums) override;
  std::vector<ErrorCode> batch_put_cancel(const std::vector<ObjectKey>& keys) override;
  std::vector<ErrorCode> batch_remove_object(const std::vector<ObjectKey>& keys) override;
  Result<std::vector<MemoryPool>> get_memory_pools() override;
  Result<std::string> client_register(const std::string& node_id) override;
  Result<ViewVersionId> client_ping(const std::string& client_id) override;
  ErrorCode register_worker(const WorkerRecord& rec) override;
  ErrorCode register_memory_pool(const MemoryPool& pool) override;
  ErrorCode worker_heartbeat(const WorkerId& id) override;

 private:
  Result<std::string> call(uint32_t method, const std::string& req);
  std::shared_ptr<net::RpcClient> current(uint64_t* gen) const;
  // Connects to the next reachable endpoint after the active one unless another thread already did (gen moved on).
  void rotate(uint64_t seen_gen);
  mutable std::mutex ep_mu_;
  std::vector<std::pair<std::string, uint16_t>> endpoints_;
  size_t active_ = 0;
  uint64_t gen_ = 0;
  std::shared_ptr<net::RpcClient> rpc_;
  int timeout_ms_ = 30000;
  int connect_timeout_ms_ = 3000;
  int failover_budget_ms_ = 15000;
  std::atomic<uint64_t> failovers_{0};
};

struct KeystoneBundle {
  std::shared_ptr<coord::CoordService> coord;
  std::shared_ptr<keystone::KeystoneService> keystone;
  std::unique_ptr<RpcService> rpc;
};
// Builds coordination client + KeystoneService + RpcService from a config and starts them
// (reference create_and_start_keystone, rpc_service.cpp:434-467).
Result<KeystoneBundle> create_and_start_keystone(const KeystoneConfig& config);

}  // namespace bb::rpc
