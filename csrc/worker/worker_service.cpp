#include "common/fault.h"
#include "worker/worker_service.h"

#include "common/audit.h"
#include "common/tenant.h"

#include <thread>

#include <chrono>
#include <stdexcept>

#include "common/checksum.h"
#include "common/log.h"
#include "common/yaml.h"
#include "rpc/wire.h"

namespace bb::worker {

// ================================================================ config
Result<WorkerServiceConfig> worker_config_from_json(const Json& root, std::string* err) {
  auto fail = [&](ErrorCode ec, const std::string& m) -> Result<WorkerServiceConfig> {
    if (err) *err = m;
    return ec;
  };
  if (!root.is_object() || !root.at("worker").is_object()) return fail(ErrorCode::MISSING_REQUIRED_FIELD, "missing top-level 'worker' section");
  const Json& w = root.at("worker");
  WorkerServiceConfig c;
  c.worker_id = w.at("worker_id").as_string();
  c.node_id = w.at("node_id").as_string();
  if (w.contains("cluster_id")) c.cluster_id = w.at("cluster_id").as_string();
  const Json& ee = w.contains("coord_endpoints") ? w.at("coord_endpoints") : w.at("etcd_endpoints");
  if (ee.is_array()) {
    for (const auto& e : ee.as_array()) c.etcd_endpoints += (c.etcd_endpoints.empty() ? "" : ",") + e.as_string();
  } else if (!ee.is_null()) {
    c.etcd_endpoints = ee.as_string();
  }
  if (w.contains("keystone_address")) c.keystone_address = w.at("keystone_address").as_string();
  if (w.contains("rpc_endpoint")) c.rpc_endpoint = w.at("rpc_endpoint").as_string();
  if (w.contains("http_metrics_port")) c.http_metrics_port = static_cast<int>(w.at("http_metrics_port").as_int(-1));
  if (w.contains("auth_token")) c.auth_token = w.at("auth_token").as_string();
  if (w.contains("encrypt_transport")) c.encrypt_transport = w.at("encrypt_transport").as_bool();
  if (w.contains("auth_token_ro")) c.auth_token_ro = w.at("auth_token_ro").as_string();
  if (w.contains("tenants_file")) c.tenants_file = w.at("tenants_file").as_string();
  if (w.contains("http_auth_token")) c.http_auth_token = w.at("http_auth_token").as_string();
  if (w.contains("audit_log")) c.audit_log = w.at("audit_log").as_string();
  if (w.contains("at_rest_key")) c.at_rest_key = w.at("at_rest_key").as_string();
  if (w.contains("ucx_endpoint")) c.ucx_endpoint = w.at("ucx_endpoint").as_string();
  if (w.contains("data_endpoint")) c.ucx_endpoint = w.at("data_endpoint").as_string();
  if (w.contains("interconnects")) {
    c.interconnects.clear();
    for (const auto& e : w.at("interconnects").as_array()) c.interconnects.push_back(e.as_string());
  }
  c.max_bw_gbps = w.at("max_bw_gbps").as_double(0.0);
  c.numa_node = static_cast<int>(w.at("numa_node").as_int(-1));
  if (w.contains("version")) c.version = w.at("version").as_string();
  c.lease_ttl_sec = w.at("lease_ttl_sec").as_int(c.lease_ttl_sec);
  c.heartbeat_interval_sec = w.at("heartbeat_interval_sec").as_int(c.heartbeat_interval_sec);
  c.allocation_poll_interval_ms = w.at("allocation_poll_interval_ms").as_int(c.allocation_poll_interval_ms);
  if (w.contains("fabric_domain")) c.fabric_domain = w.at("fabric_domain").as_string();
  if (w.contains("listen_address") && !w.contains("data_endpoint") && !w.contains("ucx_endpoint")) c.ucx_endpoint = w.at("listen_address").as_string();
  if (w.at("transport").is_object()) {
    c.transport = CxlTransportConfig::from_json(w.at("transport"));
    c.has_transport = true;
  }
  if (w.at("allocation").is_object()) c.preferred_tiers = tier_rules_from_json(w.at("allocation").at("preferred_tiers"));
  if (c.worker_id.empty() && !c.node_id.empty()) c.worker_id = c.node_id;  // cxl_worker.yaml names only the node
  if (c.worker_id.empty()) return fail(ErrorCode::MISSING_REQUIRED_FIELD, "worker.worker_id is required");
  if (c.node_id.empty()) c.node_id = c.worker_id;
  if (c.lease_ttl_sec <= 0 || c.heartbeat_interval_sec <= 0) return fail(ErrorCode::VALUE_OUT_OF_RANGE, "worker lease/heartbeat must be positive");
  if (c.heartbeat_interval_sec >= c.lease_ttl_sec) return fail(ErrorCode::INVALID_CONFIGURATION, "heartbeat_interval_sec must be shorter than lease_ttl_sec");
  // pools may live at top level (reference worker.yaml) or under worker: (reference cxl_worker.yaml)
  const Json& pools = root.contains("storage_pools") ? root.at("storage_pools") : w.at("storage_pools");
  for (const auto& p : pools.as_array()) {
    StoragePoolConfig pc;
    pc.pool_id = p.at("pool_id").as_string();
    auto sc = parse_storage_class(p.at("storage_class").as_string());
    if (pc.pool_id.empty() || !sc) return fail(ErrorCode::INVALID_CONFIGURATION, "storage pool needs pool_id and a valid storage_class");
    pc.storage_class = *sc;
    const Json& size = p.contains("size_bytes") ? p.at("size_bytes") : p.at("capacity");
    if (size.is_number()) pc.size_bytes = size.as_uint();
    else if (auto v = parse_size(size.as_string())) pc.size_bytes = *v;
    if (pc.size_bytes == 0) return fail(ErrorCode::VALUE_OUT_OF_RANGE, "storage pool " + pc.pool_id + " has no size");
    pc.mount_path = p.contains("mount_path") ? p.at("mount_path").as_string() : p.at("path").as_string();
    const Json& cfg = p.at("config");  // cxl_worker.yaml style nested block
    if (cfg.is_object()) {
      pc.cxl = CxlMemoryPoolConfig::from_json(cfg);
      pc.cxl.capacity = pc.size_bytes;
      if (!pc.cxl.dax_device.empty()) pc.mount_path = pc.cxl.dax_device;
      if (pc.cxl.enable_numa_binding || !cfg.contains("enable_numa_binding")) pc.numa_node = pc.cxl.numa_node;
    }
    pc.gpu_device_id = static_cast<int>(p.at("gpu_device_id").as_int(0));
    if (p.contains("numa_node")) pc.numa_node = static_cast<int>(p.at("numa_node").as_int(-1));
    pc.queue_depth = static_cast<uint32_t>(p.at("queue_depth").as_int(64));
    pc.pin_memory = p.at("pin_memory").as_bool(false);
    pc.shared_memory = p.at("shared_memory").as_bool(false);
    pc.encrypt_at_rest = p.at("encrypt_at_rest").as_bool(false);
    c.storage_pools.push_back(std::move(pc));
  }
  return c;
}

WorkerServiceConfig load_worker_config_from_file(const std::string& path) {
  std::string err;
  auto root = load_yaml_file(path, &err);
  if (!root) throw std::runtime_error("Failed to load worker config '" + path + "': " + err);
  auto c = worker_config_from_json(*root, &err);
  if (!c.ok()) throw std::runtime_error("Invalid worker config '" + path + "': " + err);
  return c.value();
}

// ================================================================ service
WorkerService::WorkerService(const WorkerServiceConfig& config, std::shared_ptr<coord::CoordService> coord,
                             std::shared_ptr<rpc::KeystoneApi> keystone)
    : config_(config), coord_(std::move(coord)), keystone_(std::move(keystone)) {
  register_data_handlers();
}

WorkerService::~WorkerService() { stop(); }

ErrorCode WorkerService::add_storage_pool(const std::string& pool_id, std::unique_ptr<StorageBackend> backend) {
  if (!backend || pool_id.empty()) return ErrorCode::INVALID_PARAMETERS;
  std::lock_guard<std::mutex> lk(pools_mu_);
  if (pools_.count(pool_id)) return ErrorCode::MEMORY_POOL_ALREADY_EXISTS;
  backend->set_pool_id(pool_id);
  pools_[pool_id] = std::move(backend);
  return ErrorCode::OK;
}

ErrorCode WorkerService::create_storage_pools_from_config() {
  for (const auto& pc : config_.storage_pools) {
    BackendOptions o;
    o.mount_path = pc.mount_path;
    o.gpu_device_id = pc.gpu_device_id;
    o.numa_node = pc.numa_node >= 0 ? pc.numa_node : config_.numa_node;
    o.queue_depth = pc.queue_depth;
    o.pin_memory = pc.pin_memory;
    o.shared_memory = pc.shared_memory;
    if (pc.encrypt_at_rest) {
      o.at_rest_scope = pc.pool_id;
      o.at_rest_key = config_.at_rest_key;
      if (o.at_rest_key.empty())
        if (const char* e = std::getenv("BB_AT_REST_KEY")) o.at_rest_key = e;
      if (o.at_rest_key.empty()) {
        BB_LOG(ERROR) << "pool " << pc.pool_id << ": encrypt_at_rest needs a key (worker at_rest_key: or BB_AT_REST_KEY)";
        return ErrorCode::CONFIG_ERROR;
      }
    }
    o.interleave_granularity = pc.cxl.interleave_granularity ? pc.cxl.interleave_granularity : 256;
    o.persistent = pc.cxl.is_persistent;
    auto b = create_storage_backend(pc.storage_class, pc.size_bytes, o);
    if (!b) {
      BB_LOG(ERROR) << "worker " << config_.worker_id << ": no backend for pool " << pc.pool_id << " (" << to_string(pc.storage_class) << ")";
      return ErrorCode::ALLOCATION_FAILED;
    }
    {
      std::lock_guard<std::mutex> lk(pools_mu_);
      pool_cfg_[pc.pool_id] = pc;
    }
    BB_TRY(add_storage_pool(pc.pool_id, std::move(b)));
  }
  return ErrorCode::OK;
}

ErrorCode WorkerService::initialize() {
  if (initialized_.load()) return ErrorCode::OK;
  if (!coord_ && !config_.etcd_endpoints.empty() && config_.etcd_endpoints != "none") {
    coord_ = std::make_shared<coord::CoordService>(config_.etcd_endpoints);
  }
  if (coord_ && coord_->connect() != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
  if (!coord_ && !keystone_ && !config_.keystone_address.empty()) {
    auto c = std::make_shared<rpc::KeystoneRpcClient>();
    if (c->connect(config_.keystone_address) != ErrorCode::OK) return ErrorCode::CONNECTION_FAILED;
    keystone_ = c;
  }
  {
    std::lock_guard<std::mutex> lk(pools_mu_);
    for (auto& [id, b] : pools_) {
      ErrorCode ec = b->initialize();
      if (ec != ErrorCode::OK) {
        BB_LOG(ERROR) << "worker " << config_.worker_id << ": pool " << id << " failed to initialise: " << to_string(ec);
        return ErrorCode::INITIALIZATION_FAILED;
      }
    }
  }
  auto hp = split_host_port(config_.ucx_endpoint);
  if (!hp) return ErrorCode::INVALID_ADDRESS;
  if (!config_.auth_token.empty()) net::set_cluster_token(config_.auth_token);
  if (config_.encrypt_transport) net::set_transport_encryption(true);
  if (!config_.auth_token_ro.empty()) net::set_cluster_token_ro(config_.auth_token_ro);
  if (!config_.http_auth_token.empty()) net::set_http_token(config_.http_auth_token);
  if (!config_.audit_log.empty() && !audit::open(config_.audit_log)) {
    BB_LOG(ERROR) << "worker " << config_.worker_id << ": cannot open the audit log " << config_.audit_log;
    return ErrorCode::INVALID_CONFIGURATION;
  }
  if (!config_.tenants_file.empty()) {
    std::string err;
    if (load_tenants_file(config_.tenants_file, &err) != ErrorCode::OK) {
      BB_LOG(ERROR) << "worker " << config_.worker_id << ": " << err;
      return ErrorCode::INVALID_CONFIGURATION;
    }
  } else {
    reload_tenants_if_changed();  // BB_TENANTS_FILE
  }
  data_server_.set_socket_buffers(4 << 20);  // bulk transfers: fewer wake-ups per megabyte
  const unsigned hw = std::thread::hardware_concurrency();
  ErrorCode ec = data_server_.start(hp->first, static_cast<uint16_t>(hp->second), static_cast<int>(std::min(16u, std::max(4u, hw / 2))));
  if (ec != ErrorCode::OK) return ec;
  if (config_.http_metrics_port >= 0) {
    http_server_.route("/metrics", [this](const std::string&, const std::string&) {
      return net::HttpResponse{200, "text/plain; version=0.0.4; charset=utf-8", metrics_text()};
    });
    http_server_.route("/healthz", [this](const std::string&, const std::string&) {
      return running_.load() ? net::HttpResponse{200, "text/plain; charset=utf-8", "ok\n"}
                             : net::HttpResponse{503, "text/plain; charset=utf-8", "not registered\n"};
    });
    http_server_.route("/stats", [this](const std::string&, const std::string&) {
      return net::HttpResponse{200, "application/json", get_stats().dump()};
    });
    if (http_server_.start(hp->first, static_cast<uint16_t>(config_.http_metrics_port), 1) != ErrorCode::OK)
      BB_LOG(WARNING) << "worker " << config_.worker_id << ": metrics endpoint on port " << config_.http_metrics_port << " unavailable";
  }
  initialized_.store(true);
  return ErrorCode::OK;
}

std::string WorkerService::data_endpoint() const {
  auto hp = split_host_port(config_.ucx_endpoint);
  std::string host = hp ? hp->first : "127.0.0.1";
  if (host == "0.0.0.0" || host.empty()) host = "127.0.0.1";
  return host + ":" + std::to_string(data_server_.port());
}

void WorkerService::set_pool_rkey_hex(const std::string& pool_id, const std::string& hex) {
  std::lock_guard<std::mutex> lk(pools_mu_);
  pool_rkey_hex_[pool_id] = hex;
}

MemoryPool WorkerService::describe_pool(const std::string& pool_id, const StorageBackend& b) const {
  MemoryPool p;
  p.id = pool_id;
  p.node_id = config_.node_id;
  p.worker_id = config_.worker_id;
  p.base_addr = b.get_base_address();
  p.size = b.get_total_capacity();
  p.used = b.get_used_capacity();
  p.storage_class = b.get_storage_class();
  p.ucx_endpoint = data_endpoint();
  p.ucx_remote_addr = b.get_base_address();
  auto rk = pool_rkey_hex_.find(pool_id);
  p.ucx_rkey_hex = rk != pool_rkey_hex_.end() ? rk->second : b.registration_key_hex();
  auto pc = pool_cfg_.find(pool_id);
  p.gpu_device_id = b.get_storage_class() == StorageClass::RAM_GPU ? (pc != pool_cfg_.end() ? pc->second.gpu_device_id : 0) : -1;
  p.numa_node = pc != pool_cfg_.end() && pc->second.numa_node >= 0 ? pc->second.numa_node : config_.numa_node;
  p.max_bw_gbps = config_.max_bw_gbps;
  p.fabric_domain = config_.fabric_domain;
  if (pc != pool_cfg_.end()) p.mount_path = pc->second.mount_path;
  return p;
}

std::vector<MemoryPool> WorkerService::advertised_pools() const {
  std::vector<MemoryPool> v;
  std::lock_guard<std::mutex> lk(pools_mu_);
  for (const auto& [id, b] : pools_) v.push_back(describe_pool(id, *b));
  return v;
}

ErrorCode WorkerService::register_all() {
  WorkerRecord rec;
  rec.worker_id = config_.worker_id;
  rec.node_id = config_.node_id;
  rec.rpc_endpoint = config_.rpc_endpoint;
  rec.ucx_endpoint = data_endpoint();
  rec.interconnects = config_.interconnects;
  if (config_.has_transport) {
    bool cxl_present = false, have_gpu = false;
    {
      std::lock_guard<std::mutex> lk(pools_mu_);
      for (const auto& [id, b] : pools_) {
        const StorageClass sc = b->get_storage_class();
        have_gpu |= sc == StorageClass::RAM_GPU;
        if (auto* cx = dynamic_cast<CxlMemoryBackend*>(b.get())) cxl_present |= cx->is_dax();
      }
    }
    rec.interconnects = config_.transport.resolve_interconnects(cxl_present, have_gpu);
  }
  rec.max_bw_gbps = config_.max_bw_gbps;
  rec.numa_node = config_.numa_node;
  rec.version = config_.version;
  const auto pools = advertised_pools();
  for (const auto& p : pools)
    if (std::find(rec.storage_classes.begin(), rec.storage_classes.end(), p.storage_class) == rec.storage_classes.end())
      rec.storage_classes.push_back(p.storage_class);
  if (coord_) {
    const std::string base = cluster_prefix() + "workers/" + config_.worker_id;
    if (coord_->put(base, to_json(rec).dump()) != ErrorCode::OK) return ErrorCode::SERVICE_REGISTRATION_FAILED;
    for (const auto& p : pools)
      if (coord_->put(base + "/memory_pools/" + p.id, to_json(p).dump()) != ErrorCode::OK) return ErrorCode::SERVICE_REGISTRATION_FAILED;
    if (coord_->put_with_ttl(cluster_prefix() + "heartbeat/" + config_.worker_id, std::to_string(std::time(nullptr)), config_.lease_ttl_sec) != ErrorCode::OK)
      return ErrorCode::SERVICE_REGISTRATION_FAILED;
  } else if (keystone_) {
    BB_TRY(keystone_->register_worker(rec));
    for (const auto& p : pools) BB_TRY(keystone_->register_memory_pool(p));
  }
  return ErrorCode::OK;
}

ErrorCode WorkerService::start() {
  if (!initialized_.load()) return ErrorCode::INVALID_STATE;
  if (running_.exchange(true)) return ErrorCode::INVALID_STATE;
  ErrorCode ec = register_all();
  if (ec != ErrorCode::OK) {
    running_.store(false);
    return ec;
  }
  heartbeat_thread_ = std::thread([this] { heartbeat_loop(); });
  reaper_thread_ = std::thread([this] { reservation_reaper_loop(); });
  return ErrorCode::OK;
}

// Reservation sweep: a writer that got placements (put_start -> D_RESERVE) and vanished never commits; its tokens run
// out, the backend takes the ranges back, and the worker tells the Keystone through the coordination store -- the only
// channel workers and Keystone share, as in the reference -- which drops the PENDING object and its
// ledger entry.  No Keystone-side GC pass is involved.
size_t WorkerService::reap_reservations_once() {
  std::vector<std::pair<std::string, std::string>> expired;  // (token id, object key)
  {
    std::lock_guard<std::mutex> lk(pools_mu_);
    for (auto& [id, b] : pools_) {
      auto v = b->reap_expired_reservations();
      expired.insert(expired.end(), v.begin(), v.end());
    }
  }
  if (expired.empty()) return 0;
  reservations_expired_ += expired.size();
  if (coord_ && coord_->is_connected()) {
    for (const auto& [token, owner] : expired) {
      const ErrorCode ec = coord_->put_with_ttl(cluster_prefix() + "reservations_expired/" + config_.worker_id + "/" + token, owner, 120);
      if (ec != ErrorCode::OK) BB_LOG(WARNING) << "worker " << config_.worker_id << ": cannot report expired reservation " << token;
    }
  }
  BB_LOG(INFO) << "worker " << config_.worker_id << ": reclaimed " << expired.size() << " expired shard reservations";
  return expired.size();
}

void WorkerService::reservation_reaper_loop() {
  const auto period = std::chrono::milliseconds(std::max<int64_t>(10, config_.allocation_poll_interval_ms));
  while (true) {
    {
      std::unique_lock<std::mutex> lk(sleep_mu_);
      sleep_cv_.wait_for(lk, period, [this] { return !running_.load(); });
    }
    if (!running_.load()) return;
    reap_reservations_once();
  }
}

void WorkerService::heartbeat_loop() {
  while (true) {
    {
      std::unique_lock<std::mutex> lk(sleep_mu_);
      sleep_cv_.wait_for(lk, std::chrono::seconds(config_.heartbeat_interval_sec), [this] { return !running_.load(); });
    }
    if (!running_.load()) return;
    reload_tenants_if_changed();  // an edited tenant table takes effect without a restart
    if (drop_heartbeat_.load() || fault::fire("drop_heartbeat")) continue;
    ErrorCode ec = ErrorCode::OK;
    if (coord_) ec = coord_->put_with_ttl(cluster_prefix() + "heartbeat/" + config_.worker_id, std::to_string(std::time(nullptr)), config_.lease_ttl_sec);
    else if (keystone_) ec = keystone_->worker_heartbeat(config_.worker_id);
    if (ec == ErrorCode::INVALID_WORKER) ec = register_all();  // keystone forgot us (restart / takeover)
    if (ec != ErrorCode::OK) BB_LOG(WARNING) << "worker " << config_.worker_id << ": heartbeat failed: " << to_string(ec);
    else heartbeats_sent_++;
  }
}

void WorkerService::stop() {
  if (running_.exchange(false)) {
    {
      std::lock_guard<std::mutex> lk(sleep_mu_);
      sleep_cv_.notify_all();
    }
    if (heartbeat_thread_.joinable()) heartbeat_thread_.join();
    if (reaper_thread_.joinable()) reaper_thread_.join();
    if (coord_ && coord_->is_connected()) {
      const std::string base = cluster_prefix() + "workers/" + config_.worker_id;
      if (auto st = coord_->store()) st->del_prefix(base + "/");
      coord_->del(base);
      coord_->del(cluster_prefix() + "heartbeat/" + config_.worker_id);
    }
  }
  if (initialized_.exchange(false)) {
    http_server_.stop();
    data_server_.stop();
    std::lock_guard<std::mutex> lk(pools_mu_);
    for (auto& [id, b] : pools_) b->shutdown();
  }
}

void WorkerService::inject_fault(const std::string& fault) { drop_heartbeat_.store(fault == "drop_heartbeat"); }

StorageBackend* WorkerService::backend(const std::string& pool_id) {
  std::lock_guard<std::mutex> lk(pools_mu_);
  auto it = pools_.find(pool_id);
  return it == pools_.end() ? nullptr : it->second.get();
}

Json WorkerService::get_stats() const {
  Json j = Json::object();
  j["worker_id"] = config_.worker_id;
  j["node_id"] = config_.node_id;
  j["running"] = running_.load();
  j["data_endpoint"] = data_endpoint();
  j["heartbeats_sent"] = heartbeats_sent_.load();
  j["requests_served"] = data_server_.requests_served();
  Json pools = Json::array();
  std::lock_guard<std::mutex> lk(pools_mu_);
  for (const auto& [id, b] : pools_) {
    const StorageStats s = b->get_stats();
    Json p = Json::object();
    p["pool_id"] = id;
    p["storage_class"] = std::string(to_string(b->get_storage_class()));
    p["total_capacity"] = s.total_capacity;
    p["used_capacity"] = s.used_capacity;
    p["available_capacity"] = s.available_capacity;
    p["utilization"] = s.utilization;
    p["fragmentation"] = s.fragmentation;
    p["num_reservations"] = s.num_reservations;
    p["num_committed_shards"] = s.num_committed_shards;
    p["bytes_written"] = s.bytes_written;
    p["bytes_read"] = s.bytes_read;
    p["io_errors"] = s.io_errors;
    pools.push_back(p);
  }
  j["pools"] = pools;
  return j;
}

std::string WorkerService::metrics_text() const {
  std::string out;
  auto esc = [](const std::string& v) {
    std::string r;
    for (char c : v) {
      if (c == '\\' || c == '"') r.push_back('\\');
      r.push_back(c == '\n' ? ' ' : c);
    }
    return r;
  };
  auto family = [&](const char* name, const char* type, const char* help) {
    out += "# HELP ";
    out += name;
    out += ' ';
    out += help;
    out += "\n# TYPE ";
    out += name;
    out += ' ';
    out += type;
    out += '\n';
  };
  const std::string wl = "worker=\"" + esc(config_.worker_id) + "\",node=\"" + esc(config_.node_id) + "\"";
  family("bb_worker_up", "gauge", "1 while the worker is registered and heartbeating");
  out += "bb_worker_up{" + wl + "} " + std::string(running_.load() ? "1" : "0") + "\n";
  family("bb_worker_heartbeats_total", "counter", "lease refreshes sent");
  out += "bb_worker_heartbeats_total{" + wl + "} " + std::to_string(heartbeats_sent_.load()) + "\n";
  family("bb_worker_data_requests_total", "counter", "requests served by the data server");
  out += "bb_worker_data_requests_total{" + wl + "} " + std::to_string(data_server_.requests_served()) + "\n";
  family("bb_worker_data_shm_requests_total", "counter", "data-server requests served over same-host shared-memory channels");
  out += "bb_worker_data_shm_requests_total{" + wl + "} " + std::to_string(data_server_.shm_requests_served()) + "\n";
  family("bb_worker_data_secure_handshakes_total", "counter", "data-server connections that switched to sealed frames (encrypt_transport)");
  out += "bb_worker_data_secure_handshakes_total{" + wl + "} " + std::to_string(data_server_.secure_handshakes()) + "\n";
  family("bb_worker_data_auth_failures_total", "counter", "denied handshakes and frames that failed authentication at the data server");
  out += "bb_worker_data_auth_failures_total{" + wl + "} " + std::to_string(data_server_.auth_failures()) + "\n";
  family("bb_worker_data_read_only_denials_total", "counter", "data-server requests of read-only members outside the read-only list");
  out += "bb_worker_data_read_only_denials_total{" + wl + "} " + std::to_string(data_server_.read_only_denials()) + "\n";
  family("bb_worker_data_connections", "gauge", "open data-server connections");
  out += "bb_worker_data_connections{" + wl + "} " + std::to_string(data_server_.connection_count()) + "\n";
  struct Row {
    std::string labels;
    StorageStats s;
    uint64_t device_copies;
  };
  std::vector<Row> rows;
  {
    std::lock_guard<std::mutex> lk(pools_mu_);
    for (const auto& [id, b] : pools_)
      rows.push_back({wl + ",pool=\"" + esc(id) + "\",tier=\"" + std::string(to_string(b->get_storage_class())) + "\"", b->get_stats(), b->device_copies()});
  }
  auto series = [&](const char* name, const char* type, const char* help, auto value) {
    family(name, type, help);
    for (const Row& r : rows) {
      out += name;
      out += '{' + r.labels + "} " + value(r) + "\n";
    }
  };
  series("bb_pool_capacity_bytes", "gauge", "pool capacity", [](const Row& r) { return std::to_string(r.s.total_capacity); });
  // placement is decided by the Keystone's allocator (its /metrics has the per-tier used bytes); this is the part
  // handed out through the worker-side reserve / commit protocol
  series("bb_pool_reserved_bytes", "gauge", "bytes reserved or committed through the worker-side reservation protocol",
         [](const Row& r) { return std::to_string(r.s.used_capacity); });
  series("bb_pool_reservations", "gauge", "uncommitted reservations", [](const Row& r) { return std::to_string(r.s.num_reservations); });
  series("bb_pool_committed_shards", "gauge", "committed shards", [](const Row& r) { return std::to_string(r.s.num_committed_shards); });
  series("bb_pool_fragmentation_ratio", "gauge", "1 - largest free block / free bytes", [](const Row& r) { return std::to_string(r.s.fragmentation); });
  series("bb_pool_bytes_written_total", "counter", "bytes written into the pool", [](const Row& r) { return std::to_string(r.s.bytes_written); });
  series("bb_pool_bytes_read_total", "counter", "bytes read out of the pool", [](const Row& r) { return std::to_string(r.s.bytes_read); });
  series("bb_pool_io_errors_total", "counter", "failed reads / writes", [](const Row& r) { return std::to_string(r.s.io_errors); });
  series("bb_pool_fused_tier_moves_total", "counter", "tier moves executed as one fused-kernel launch", [](const Row& r) { return std::to_string(r.device_copies); });
  return out;
}

// ================================================================ data server
namespace {
// Bit 63 marks an absolute address (MemoryLocation::remote_addr); otherwise a pool offset.
uint64_t resolve_offset(const StorageBackend& b, uint64_t raw) {
  if (raw >> 63) {
    const uint64_t addr = raw & ~(1ull << 63);
    return addr >= b.get_base_address() ? addr - b.get_base_address() : ~0ull;
  }
  return raw;
}
// Overflow-safe "[off, off + len) lies inside the pool".
bool range_ok(const StorageBackend& b, uint64_t off, uint64_t len) {
  const uint64_t cap = b.get_total_capacity();
  return off != ~0ull && len <= cap && off <= cap - len;
}
// Unfinalised tile sum (BBH64 / XXH3) of a chunk that starts `pos` bytes into the hashed object (pos is tile aligned).
uint64_t chunk_tile_sum(ChecksumAlgo algo, const uint8_t* data, uint64_t n, uint64_t pos) { return tile_sum_chunk(algo, data, n, pos / 16384); }
}  // namespace

void WorkerService::register_data_handlers() {
  // Tenants (common/tenant.h) move the bytes of the placements the Keystone gave them -- and nothing else: copies between
  // pools, pulls, reservations and frees are the Keystone's and the members' business.  (Which shard a tenant may touch is
  // decided where the placement is handed out: like an rkey in the reference, a placement is the capability.)
  data_server_.allow_tenants({D_WRITE, D_READ, D_CHECKSUM, D_STATS});
  data_server_.allow_read_only({D_READ, D_CHECKSUM, D_STATS});  // read-only members read shards; they never write, copy, pull or reserve
  using C = const net::ConnPtr&;
  using S = const std::string&;
  using V = std::string_view;
  using Reply = net::RpcServer::Reply;
  // D_WRITE / D_READ are bulk methods: the request payload is consumed in place from the connection buffer and a
  // read of a host-mapped tier is answered straight out of the pool (gathered send), so a shard crosses the worker
  // with one copy on the way in and none on the way out.
  data_server_.register_view_method(D_WRITE, [this](C, V q) {
    wire::Reader r(q.data(), q.size());
    const std::string pool = r.str();
    const uint64_t off = r.u64();
    const uint32_t len = r.u32();
    wire::Writer w;
    StorageBackend* b = backend(pool);
    if (!r.ok() || q.size() < len || !b) {
      w.ec(!b ? ErrorCode::MEMORY_POOL_NOT_FOUND : ErrorCode::INVALID_PARAMETERS);
      return Reply{w.take()};
    }
    const char* payload = q.data() + (q.size() - len);
    if (fault::fire("fail_data_write")) {
      w.ec(ErrorCode::IO_ERROR);
      return Reply{w.take()};
    }
    ErrorCode wec = b->write(resolve_offset(*b, off), payload, len);
    if (wec == ErrorCode::OK && len && fault::fire("corrupt_write")) {  // silent corruption: the checksum must catch it
      const char bad = static_cast<char>(payload[len / 2] ^ 0x5A);
      b->write(resolve_offset(*b, off) + len / 2, &bad, 1);
    }
    w.ec(wec);
    return Reply{w.take()};
  });
  data_server_.register_view_method(D_READ, [this](C, V q) {
    wire::Reader r(q.data(), q.size());
    const std::string pool = r.str();
    const uint64_t off = r.u64();
    const uint32_t len = r.u32();
    StorageBackend* b = backend(pool);
    Reply rep;
    ErrorCode ec = !b ? ErrorCode::MEMORY_POOL_NOT_FOUND : !r.ok() ? ErrorCode::INVALID_PARAMETERS
                   : fault::fire("fail_data_read") ? ErrorCode::IO_ERROR : ErrorCode::OK;
    if (ec == ErrorCode::OK) {
      const uint64_t o = resolve_offset(*b, off);
      // the length comes off the wire: check it against the pool BEFORE anything is sized from it (a 4 GiB `len` used to
      // zero-fill a 4 GiB reply buffer first and fail the range check afterwards)
      const bool in_range = o != ~0ull && o + len >= o && o + len <= b->get_total_capacity();
      const void* direct = (in_range && b->get_storage_class() != StorageClass::RAM_GPU) ? b->direct_ptr(o) : nullptr;
      if (!in_range) {
        ec = ErrorCode::INVALID_PARAMETERS;
      } else if (direct && len >= 4096) {  // DRAM / CXL / mmap tiers: send from the pool itself
        b->note_read(len);
        rep.ext = direct;
        rep.ext_len = len;
        rep.head.assign(4, '\0');
      } else {
        rep.head.assign(4 + static_cast<size_t>(len), '\0');
        ec = b->read(o, rep.head.data() + 4, len);
      }
    }
    if (ec != ErrorCode::OK) {
      rep.head.assign(4, '\0');
      rep.ext = nullptr;
      rep.ext_len = 0;
    }
    const uint32_t e = static_cast<uint32_t>(ec);
    std::memcpy(rep.head.data(), &e, 4);
    return rep;
  });
  data_server_.register_method(D_CHECKSUM, [this](C, S q) {
    wire::Reader r(q);
    const std::string pool = r.str();
    const uint64_t off = r.u64();
    const uint64_t len = r.u64();
    const auto algo = static_cast<ChecksumAlgo>(r.u32());
    wire::Writer w;
    StorageBackend* b = backend(pool);
    if (!b || !r.ok()) {
      w.ec(!b ? ErrorCode::MEMORY_POOL_NOT_FOUND : ErrorCode::INVALID_PARAMETERS);
      return w.take();
    }
    // The 64-bit length comes off the wire: validate it against the pool before anything is sized from it, and
    // hash in bounded chunks (CRC32C streams; BBH64 tile sums are additive) instead of buffering the shard.
    const uint64_t o = resolve_offset(*b, off);
    if (!range_ok(*b, o, len)) {
      w.ec(ErrorCode::MEMORY_ACCESS_ERROR);
      return w.take();
    }
    constexpr uint64_t kChunk = 8ull << 20;  // a multiple of the BBH64 tile
    std::vector<uint8_t> buf(std::min(len, kChunk));
    uint32_t crc = 0;
    uint64_t tile_sum = 0;
    ErrorCode ec = ErrorCode::OK;
    for (uint64_t pos = 0; pos < len && ec == ErrorCode::OK; pos += kChunk) {
      const uint64_t n = std::min(kChunk, len - pos);
      ec = b->read(o + pos, buf.data(), n);
      if (ec != ErrorCode::OK) break;
      if (algo == ChecksumAlgo::CRC32C) crc = crc32c(buf.data(), n, crc);
      else if (is_tile_sum(algo)) tile_sum += chunk_tile_sum(algo, buf.data(), n, pos);
    }
    w.ec(ec);
    if (ec == ErrorCode::OK)
      w.u64(algo == ChecksumAlgo::CRC32C ? crc : is_tile_sum(algo) ? tile_sum_finalize(tile_sum, len) : 0);
    return w.take();
  });
  // Intra-worker tier move (GPU slab -> DRAM -> NVMe ...): bytes stay inside this process.
  data_server_.register_method(D_COPY, [this](C, S q) {
    wire::Reader r(q);
    const std::string sp = r.str();
    const uint64_t so = r.u64();
    const std::string dp = r.str();
    const uint64_t doff = r.u64();
    const uint64_t len = r.u64();
    const auto algo = static_cast<ChecksumAlgo>(r.u32());
    wire::Writer w;
    StorageBackend* sb = backend(sp);
    StorageBackend* db = backend(dp);
    if (!r.ok() || !sb || !db) {
      w.ec(!r.ok() ? ErrorCode::INVALID_PARAMETERS : ErrorCode::MEMORY_POOL_NOT_FOUND);
      return w.take();
    }
    const uint64_t s0 = resolve_offset(*sb, so), d0 = resolve_offset(*db, doff);
    if (!range_ok(*sb, s0, len) || !range_ok(*db, d0, len)) {
      w.ec(ErrorCode::MEMORY_ACCESS_ERROR);
      return w.take();
    }
    // Fused-kernel tier move: when one side is the GPU slab and the other is addressable by CUDA (another
    // slab, or a pinned DRAM pool), ONE launch copies the shard and computes its digest on the tensor cores.
    {
      uint64_t digest = 0;
      ErrorCode dc = sb->device_copy(*db, true, s0, d0, len, algo, &digest);
      if (dc == ErrorCode::NOT_IMPLEMENTED) dc = db->device_copy(*sb, false, d0, s0, len, algo, &digest);
      if (dc != ErrorCode::NOT_IMPLEMENTED) {
        if (dc == ErrorCode::OK) dc = db->flush();
        w.ec(dc);
        if (dc == ErrorCode::OK) w.u64(digest);
        return w.take();
      }
    }
    constexpr uint64_t kChunk = 8ull << 20;
    std::vector<uint8_t> buf(std::min(len, kChunk));
    uint64_t tile_sum = 0;  // BBH64 tile sums are additive: the chunk (a whole number of tiles) is hashed in place
    uint32_t crc = 0;
    ErrorCode ec = ErrorCode::OK;
    for (uint64_t pos = 0; pos < len && ec == ErrorCode::OK; pos += kChunk) {
      const uint64_t n = std::min(kChunk, len - pos);
      ec = sb->read(s0 + pos, buf.data(), n);
      if (ec != ErrorCode::OK) break;
      if (algo == ChecksumAlgo::CRC32C) crc = crc32c(buf.data(), n, crc);
      else if (is_tile_sum(algo)) tile_sum += chunk_tile_sum(algo, buf.data(), n, pos);
      ec = db->write(d0 + pos, buf.data(), n);
    }
    if (ec == ErrorCode::OK) ec = db->flush();
    w.ec(ec);
    if (ec == ErrorCode::OK) w.u64(algo == ChecksumAlgo::CRC32C ? crc : is_tile_sum(algo) ? tile_sum_finalize(tile_sum, len) : 0);
    return w.take();
  });
  data_server_.register_method(D_PULL, [this](C, S q) {
    wire::Reader r(q);
    const std::string dp = r.str();
    const uint64_t doff = r.u64();
    const std::string key = r.str();  // raw registration key of the source pool (CUDA IPC handle)
    const uint64_t soff = r.u64();
    const uint64_t len = r.u64();
    const auto algo = static_cast<ChecksumAlgo>(r.u32());
    wire::Writer w;
    StorageBackend* db = backend(dp);
    if (!r.ok() || !db) {
      w.ec(!r.ok() ? ErrorCode::INVALID_PARAMETERS : ErrorCode::MEMORY_POOL_NOT_FOUND);
      return w.take();
    }
    uint64_t digest = 0;
    ErrorCode ec = db->pull_from_peer(std::vector<uint8_t>(key.begin(), key.end()), soff, resolve_offset(*db, doff), len, algo, &digest);
    w.ec(ec);
    if (ec == ErrorCode::OK) w.u64(digest);
    return w.take();
  });
  // ---- reservation protocol (Keystone -> worker).  Requests carry one pool and a list of shards / tokens.
  data_server_.register_method(D_RESERVE, [this](C, S q) {
    wire::Reader r(q);
    const std::string pool = r.str(), owner = r.str();
    const uint64_t ttl_ms = r.u64();
    const uint32_t n = r.u32();
    wire::Writer w;
    StorageBackend* b = backend(pool);
    if (!r.ok() || !b || n > 65536) {
      w.ec(!b ? ErrorCode::MEMORY_POOL_NOT_FOUND : ErrorCode::INVALID_PARAMETERS);
      return w.take();
    }
    std::vector<std::string> tokens;
    ErrorCode ec = ErrorCode::OK;
    for (uint32_t i = 0; i < n && ec == ErrorCode::OK; ++i) {
      const uint64_t off = resolve_offset(*b, r.u64());
      const uint64_t len = r.u64();
      if (!r.ok() || !range_ok(*b, off, len)) {
        ec = ErrorCode::MEMORY_ACCESS_ERROR;
        break;
      }
      auto t = b->reserve_shard_at(off, len, owner, ttl_ms);
      if (!t.ok()) ec = t.error();
      else tokens.push_back(t.value().token_id);
    }
    if (ec != ErrorCode::OK)  // all or nothing
      for (const auto& t : tokens) b->abort_shard_id(t);
    w.ec(ec);
    if (ec == ErrorCode::OK) {
      w.u32(static_cast<uint32_t>(tokens.size()));
      for (const auto& t : tokens) w.str(t);
    }
    return w.take();
  });
  auto by_token = [this](bool commit) {
    return [this, commit](C, S q) {
      wire::Reader r(q);
      const std::string pool = r.str();
      const uint32_t n = r.u32();
      wire::Writer w;
      StorageBackend* b = backend(pool);
      if (!r.ok() || !b || n > 65536) {
        w.ec(!b ? ErrorCode::MEMORY_POOL_NOT_FOUND : ErrorCode::INVALID_PARAMETERS);
        return w.take();
      }
      ErrorCode worst = ErrorCode::OK;
      for (uint32_t i = 0; i < n; ++i) {
        const std::string tok = r.str();
        if (!r.ok()) {
          worst = ErrorCode::INVALID_PARAMETERS;
          break;
        }
        const ErrorCode ec = commit ? b->commit_shard_id(tok) : b->abort_shard_id(tok);
        if (ec != ErrorCode::OK && worst == ErrorCode::OK) worst = ec;
      }
      w.ec(worst);
      return w.take();
    };
  };
  data_server_.register_method(D_COMMIT, by_token(true));
  data_server_.register_method(D_ABORT, by_token(false));
  data_server_.register_method(D_FREE, [this](C, S q) {
    wire::Reader r(q);
    const std::string pool = r.str();
    const uint32_t n = r.u32();
    wire::Writer w;
    StorageBackend* b = backend(pool);
    if (!r.ok() || !b || n > 65536) {
      w.ec(!b ? ErrorCode::MEMORY_POOL_NOT_FOUND : ErrorCode::INVALID_PARAMETERS);
      return w.take();
    }
    ErrorCode worst = ErrorCode::OK;
    for (uint32_t i = 0; i < n; ++i) {
      const uint64_t off = resolve_offset(*b, r.u64());
      const uint64_t len = r.u64();
      if (!r.ok() || off == ~0ull) {
        worst = ErrorCode::INVALID_PARAMETERS;
        break;
      }
      const ErrorCode ec = b->free_shard(b->get_base_address() + off, len);
      if (ec != ErrorCode::OK && worst == ErrorCode::OK) worst = ec;
    }
    w.ec(worst);
    return w.take();
  });
  data_server_.register_method(D_STATS, [this](C, S) {
    wire::Writer w;
    w.ec(ErrorCode::OK);
    w.str(get_stats().dump());
    return w.take();
  });
}

}  // namespace bb::worker
