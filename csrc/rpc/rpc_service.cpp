#include "rpc/rpc_service.h"

#include "client/copy_mover.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <thread>

#include "common/audit.h"
#include "common/log.h"
#include "rpc/wire.h"

namespace bb::rpc {

using keystone::PutStartItem;
using keystone::ShardChecksums;
using wire::Reader;
using wire::Writer;

namespace {
void put_sums(Writer& w, const ShardChecksums& s) {
  w.u32(static_cast<uint32_t>(s.size()));
  for (const auto& c : s) {
    w.u32(static_cast<uint32_t>(c.size()));
    for (uint64_t v : c) w.u64(v);
  }
}
void get_sums(Reader& r, ShardChecksums& s) {
  s.resize(r.count(4));
  for (auto& c : s) {
    c.resize(r.count(8));
    for (auto& v : c) v = r.u64();
  }
}
void put_keys(Writer& w, const std::vector<ObjectKey>& keys) {
  w.u32(static_cast<uint32_t>(keys.size()));
  for (const auto& k : keys) w.str(k);
}
std::vector<ObjectKey> get_keys(Reader& r) {
  std::vector<ObjectKey> v(r.count(4));
  for (auto& k : v) k = r.str();
  return v;
}
void put_copies_result(Writer& w, const Result<std::vector<CopyPlacement>>& res) {
  w.ec(res.error());
  if (res.ok()) wire::put(w, res.value());
}
Result<std::vector<CopyPlacement>> get_copies_result(Reader& r) {
  const ErrorCode ec = r.ec();
  if (ec != ErrorCode::OK) return ec;
  std::vector<CopyPlacement> v;
  wire::get(r, v);
  if (!r.ok()) return ErrorCode::RPC_FAILED;
  return v;
}
std::string ec_reply(ErrorCode ec) {
  Writer w;
  w.ec(ec);
  return w.take();
}
// Key-level ACL of the tenant on whose behalf this thread runs (net::RpcServer::dispatch sets the scope); members pass.
bool acl_refused(const char* op, const std::string& key) {
  if (audit::enabled()) audit::event("acl_denied", {{"op", op}, {"key", key}});
  return false;
}
bool may_read(const std::string& key) {
  const Tenant* t = current_tenant();
  return !t || t->may_read(key) || acl_refused("read", key);
}
bool may_write(const std::string& key) {
  const Tenant* t = current_tenant();
  return !t || t->may_write(key) || acl_refused("write", key);
}
bool may_read_all(const std::vector<ObjectKey>& keys) {
  const Tenant* t = current_tenant();
  if (!t) return true;
  for (const auto& k : keys)
    if (!t->may_read(k)) return acl_refused("read", k);
  return true;
}
bool may_write_all(const std::vector<ObjectKey>& keys) {
  const Tenant* t = current_tenant();
  if (!t) return true;
  for (const auto& k : keys)
    if (!t->may_write(k)) return acl_refused("write", k);
  return true;
}
std::string ecs_reply(const std::vector<ErrorCode>& v) {
  Writer w;
  w.ec(ErrorCode::OK);
  w.u32(static_cast<uint32_t>(v.size()));
  for (auto e : v) w.ec(e);
  return w.take();
}
}  // namespace

// ================================================================ server
RpcService::RpcService(std::shared_ptr<keystone::KeystoneService> keystone, const KeystoneConfig& config)
    : keystone_(std::move(keystone)), config_(config) {
  register_handlers();
}

RpcService::~RpcService() { stop(); }

void RpcService::leader_only(uint32_t method, net::RpcServer::Handler h) {
  auto ks = keystone_;
  rpc_.register_method(method, [ks, h = std::move(h)](const net::ConnPtr& c, const std::string& q) {
    if (!ks->is_leader()) return ec_reply(ErrorCode::NOT_LEADER);
    return h(c, q);
  });
}

void RpcService::register_handlers() {
  // what a holder of the read-only token may ask the Keystone: existence, placements (to read the data), statistics, listings
  rpc_.allow_read_only({M_OBJECT_EXISTS, M_GET_WORKERS, M_GET_CLUSTER_STATS, M_GET_VIEW_VERSION, M_BATCH_OBJECT_EXISTS, M_BATCH_GET_WORKERS,
                        M_GET_MEMORY_POOLS, M_GET_WORKERS_INFO, M_LIST_OBJECTS, M_CLIENT_REGISTER, M_CLIENT_PING, M_TENANT_USAGE});
  // what a tenant may ask (common/tenant.h): the object-level calls, each checked against its key grants below; the
  // cluster-management methods (workers, pools, migrate, drain, scrub, compact, remove_all) need an `admin` tenant or a member
  rpc_.allow_tenants({M_OBJECT_EXISTS, M_GET_WORKERS, M_PUT_START, M_PUT_COMPLETE, M_PUT_CANCEL, M_REMOVE_OBJECT, M_GET_CLUSTER_STATS,
                      M_GET_VIEW_VERSION, M_BATCH_OBJECT_EXISTS, M_BATCH_GET_WORKERS, M_BATCH_PUT_START, M_BATCH_PUT_COMPLETE, M_BATCH_PUT_CANCEL,
                      M_BATCH_REMOVE_OBJECT, M_CLIENT_REGISTER, M_CLIENT_PING, M_GET_MEMORY_POOLS, M_LIST_OBJECTS, M_TENANT_USAGE});
  auto denied = [ksm = keystone_] {
    ksm->count_acl_denial();
    return ErrorCode::ACCESS_DENIED;
  };
  // audit trail of the cluster-management calls (common/audit.h): who asked for what, and how it ended
  auto admin = [](const char* op, const std::string& arg, ErrorCode ec) {
    if (audit::enabled()) audit::event("admin", {{"op", op}, {"arg", arg}, {"result", to_string(ec)}});
    return ec;
  };
  auto ks = keystone_;
  using C = const net::ConnPtr&;
  using S = const std::string&;
  leader_only(M_OBJECT_EXISTS, [ks, denied](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    auto res = may_read(key) ? ks->object_exists(key) : Result<bool>(denied());
    Writer w;
    w.ec(res.error());
    w.boolean(res.ok() && res.value());
    return w.take();
  });
  leader_only(M_GET_WORKERS, [ks, denied](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    Writer w;
    if (!may_read(key)) put_copies_result(w, denied());
    else put_copies_result(w, ks->get_workers(key));
    return w.take();
  });
  leader_only(M_PUT_START, [ks](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    const uint64_t size = r.u64();
    WorkerConfig cfg;
    wire::get(r, cfg);
    const std::string cid = r.str(), node = r.str();
    Writer w;
    if (!r.ok()) w.ec(ErrorCode::INVALID_PARAMETERS);
    else put_copies_result(w, ks->put_start(key, size, cfg, cid, node));
    return w.take();
  });
  leader_only(M_PUT_COMPLETE, [ks, denied](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    ShardChecksums sums;
    get_sums(r, sums);
    if (!may_write(key)) return ec_reply(denied());
    return ec_reply(r.ok() ? ks->put_complete(key, sums) : ErrorCode::INVALID_PARAMETERS);
  });
  leader_only(M_PUT_CANCEL, [ks, denied](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    return ec_reply(may_write(key) ? ks->put_cancel(key) : denied());
  });
  leader_only(M_REMOVE_OBJECT, [ks, denied](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    return ec_reply(may_write(key) ? ks->remove_object(key) : denied());
  });
  leader_only(M_REMOVE_ALL_OBJECTS, [ks, admin](C, S) {
    auto res = ks->remove_all_objects();
    admin("remove_all_objects", res.ok() ? std::to_string(res.value()) : "", res.ok() ? ErrorCode::OK : res.error());
    Writer w;
    w.ec(res.error());
    w.u64(res.ok() ? res.value() : 0);
    return w.take();
  });
  rpc_.register_method(M_GET_CLUSTER_STATS, [ks](C, S) {
    auto res = ks->get_cluster_stats();
    Writer w;
    w.ec(res.error());
    if (res.ok()) wire::put(w, res.value());
    return w.take();
  });
  rpc_.register_method(M_GET_VIEW_VERSION, [ks](C, S) {
    Writer w;
    w.ec(ErrorCode::OK);
    w.i64(ks->get_view_version());
    return w.take();
  });
  leader_only(M_BATCH_OBJECT_EXISTS, [ks, denied](C, S q) {
    Reader r(q);
    const auto keys = get_keys(r);
    if (!may_read_all(keys)) return ec_reply(denied());  // one key outside the grants refuses the batch (a whole-batch error code)
    auto res = ks->batch_object_exists(keys);
    Writer w;
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(res.size()));
    for (const auto& e : res) {
      w.ec(e.error());
      w.boolean(e.ok() && e.value());
    }
    return w.take();
  });
  leader_only(M_BATCH_GET_WORKERS, [ks, denied](C, S q) {
    Reader r(q);
    const auto keys = get_keys(r);
    if (!may_read_all(keys)) return ec_reply(denied());
    auto res = ks->batch_get_workers(keys);
    Writer w;
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(res.size()));
    wire::PlacementBatchWriter pw(w);
    for (const auto& e : res) pw.put(e);
    return w.take();
  });
  leader_only(M_BATCH_PUT_START, [ks](C, S q) {
    Reader r(q);
    std::vector<PutStartItem> items(r.count(12));
    const bool one_config = r.boolean();  // the usual batch: one policy for every item, sent once
    WorkerConfig shared;
    if (one_config) wire::get(r, shared);
    for (auto& it : items) {
      it.key = r.str();
      it.size = r.u64();
      if (one_config) it.config = shared;
      else wire::get(r, it.config);
    }
    const std::string cid = r.str(), node = r.str();
    Writer w;
    if (!r.ok()) {
      w.ec(ErrorCode::INVALID_PARAMETERS);
      return w.take();
    }
    auto res = ks->batch_put_start(items, cid, node);
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(res.size()));
    wire::PlacementBatchWriter pw(w);
    for (const auto& e : res) pw.put(e);
    return w.take();
  });
  leader_only(M_BATCH_PUT_COMPLETE, [ks, denied](C, S q) {
    Reader r(q);
    auto keys = get_keys(r);
    std::vector<ShardChecksums> sums(r.count(4));
    for (auto& s : sums) get_sums(r, s);
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    if (!may_write_all(keys)) return ec_reply(denied());
    return ecs_reply(ks->batch_put_complete(keys, sums));
  });
  leader_only(M_BATCH_PUT_CANCEL, [ks, denied](C, S q) {
    Reader r(q);
    const auto keys = get_keys(r);
    if (!may_write_all(keys)) return ec_reply(denied());
    return ecs_reply(ks->batch_put_cancel(keys));
  });
  leader_only(M_BATCH_REMOVE_OBJECT, [ks, denied](C, S q) {
    Reader r(q);
    const auto keys = get_keys(r);
    if (!may_write_all(keys)) return ec_reply(denied());
    return ecs_reply(ks->batch_remove_object(keys));
  });
  leader_only(M_CLIENT_REGISTER, [ks](C, S q) {
    Reader r(q);
    auto res = ks->client_register(r.str());
    Writer w;
    w.ec(res.error());
    w.str(res.ok() ? res.value() : "");
    return w.take();
  });
  leader_only(M_CLIENT_PING, [ks](C, S q) {
    Reader r(q);
    auto res = ks->client_ping(r.str());
    Writer w;
    w.ec(res.error());
    w.i64(res.ok() ? res.value() : 0);
    return w.take();
  });
  rpc_.register_method(M_GET_MEMORY_POOLS, [ks](C, S) {
    std::vector<MemoryPool> pools;
    ks->get_memory_pools(pools);
    Writer w;
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(pools.size()));
    for (const auto& p : pools) wire::put(w, p);
    return w.take();
  });
  rpc_.register_method(M_REGISTER_WORKER, [ks](C, S q) {
    Reader r(q);
    auto j = Json::parse(r.str());
    if (!j) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    auto rec = worker_record_from_json(*j);
    return ec_reply(rec.ok() ? ks->register_worker(rec.value()) : rec.error());
  });
  rpc_.register_method(M_REGISTER_MEMORY_POOL, [ks](C, S q) {
    Reader r(q);
    MemoryPool p;
    wire::get(r, p);
    return ec_reply(r.ok() ? ks->register_memory_pool(p) : ErrorCode::INVALID_PARAMETERS);
  });
  rpc_.register_method(M_WORKER_HEARTBEAT, [ks](C, S q) {
    Reader r(q);
    return ec_reply(ks->worker_heartbeat(r.str()));
  });
  leader_only(M_MIGRATE_OBJECT, [ks, admin](C, S q) {
    Reader r(q);
    const std::string key = r.str();
    const auto target = static_cast<StorageClass>(r.u32());
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    return ec_reply(admin("migrate_object", key, ks->migrate_object(key, target)));
  });
  rpc_.register_method(M_GET_WORKERS_INFO, [ks](C, S) {
    std::vector<keystone::WorkerInfo> v;
    ks->get_workers_info(v);
    Writer w;
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(v.size()));
    const auto now = Clock::now();
    for (const auto& wi : v) {
      w.str(wi.worker_id);
      w.str(wi.node_id);
      w.str(wi.endpoint);
      w.i64(std::chrono::duration_cast<std::chrono::milliseconds>(now - wi.last_heartbeat).count());
      w.u32(static_cast<uint32_t>(wi.pools.size()));
      for (const auto& p : wi.pools) w.str(p);
    }
    return w.take();
  });
  leader_only(M_LIST_OBJECTS, [ks](C, S q) {
    Reader r(q);
    const std::string prefix = r.str();
    const uint64_t limit = r.u64();
    const std::string after = r.str();
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    if (const Tenant* t = current_tenant(); t && !t->may_list(prefix)) {  // list inside a grant
      acl_refused("list", prefix);
      return ec_reply(ErrorCode::ACCESS_DENIED);
    }
    const auto v = ks->list_objects(prefix, static_cast<size_t>(limit), after);
    Writer w;
    w.ec(ErrorCode::OK);
    w.u32(static_cast<uint32_t>(v.size()));
    for (const auto& o : v) {
      w.str(o.key);
      w.u64(o.size);
      w.u32(o.copies);
      w.u32(static_cast<uint32_t>(o.tier));
    }
    return w.take();
  });
  leader_only(M_DRAIN_WORKER, [ks, admin](C, S q) {
    Reader r(q);
    const std::string id = r.str();
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    auto res = ks->drain_worker(id);
    admin("drain_worker", id, res.ok() ? ErrorCode::OK : res.error());
    Writer w;
    w.ec(res.ok() ? ErrorCode::OK : res.error());
    w.u64(res.ok() ? res.value() : 0);
    return w.take();
  });
  leader_only(M_SCRUB, [ks, admin](C, S q) {
    Reader r(q);
    const std::string prefix = r.str();
    const uint64_t max_objects = r.u64();
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    auto res = ks->scrub(prefix, static_cast<size_t>(max_objects));
    admin("scrub", prefix, res.ok() ? ErrorCode::OK : res.error());
    Writer w;
    w.ec(res.ok() ? ErrorCode::OK : res.error());
    if (res.ok())
      for (uint64_t v : {res.value().objects, res.value().copies, res.value().corrupt, res.value().healed, res.value().unrecoverable, res.value().unreachable}) w.u64(v);
    return w.take();
  });
  leader_only(M_COMPACT_POOL, [ks, admin](C, S q) {
    Reader r(q);
    const std::string pool = r.str();
    const uint64_t max_moves = r.u64();
    if (!r.ok()) return ec_reply(ErrorCode::INVALID_PARAMETERS);
    auto res = ks->compact_pool(pool, static_cast<size_t>(max_moves));
    admin("compact_pool", pool, res.ok() ? ErrorCode::OK : res.error());
    Writer w;
    w.ec(res.ok() ? ErrorCode::OK : res.error());
    if (res.ok()) w.u64(res.value());
    return w.take();
  });
  leader_only(M_REMOVE_WORKER, [ks, admin](C, S q) {
    Reader r(q);
    const std::string id = r.str();
    return ec_reply(admin("remove_worker", id, ks->remove_worker(id)));
  });
  leader_only(M_TENANT_USAGE, [ks](C, S) {
    const Tenant* me = current_tenant();  // a tenant sees its own line; members and admins see every tenant
    Writer w;
    w.ec(ErrorCode::OK);
    std::vector<keystone::TenantUsage> v;
    for (auto& u : ks->tenant_usage())
      if (!me || me->admin || u.name == me->name) v.push_back(std::move(u));
    w.u32(static_cast<uint32_t>(v.size()));
    for (const auto& u : v) {
      w.str(u.name);
      for (uint64_t x : {u.used_bytes, u.objects, u.quota_bytes, u.max_objects}) w.u64(x);
    }
    return w.take();
  });

  http_.route("/metrics", [ks, this](const std::string&, const std::string&) {
    net::HttpResponse r;
    r.content_type = "text/plain; version=0.0.4; charset=utf-8";
    r.body = ks->metrics_text();
    // transport counters of this RPC server
    auto counter = [&](const char* name, const char* help, uint64_t v) {
      r.body += std::string("# HELP ") + name + " " + help + "\n# TYPE " + name + " counter\n" + name + " " + std::to_string(v) + "\n";
    };
    counter("bb_rpc_requests_total", "RPC requests served (TCP frames + shared-memory channel)", rpc_.requests_served());
    counter("bb_rpc_shm_requests_total", "RPC requests served over same-host shared-memory channels", rpc_.shm_requests_served());
    counter("bb_rpc_secure_handshakes_total", "connections that switched to AES-256-GCM sealed frames (encrypt_transport)", rpc_.secure_handshakes());
    counter("bb_rpc_auth_failures_total", "denied token handshakes, requests without the token, frames that failed authentication", rpc_.auth_failures());
    counter("bb_rpc_read_only_denials_total", "requests of read-only members for methods outside the read-only list", rpc_.read_only_denials());
    if (audit::enabled()) counter("bb_audit_events_total", "lines appended to the audit log (common/audit.h)", audit::events_written());
    counter("bb_rpc_tenant_handshakes_total", "connections admitted as a tenant (common/tenant.h)", rpc_.tenant_handshakes());
    counter("bb_rpc_tenant_denials_total", "tenant requests for methods outside the tenant list, or of tenants that left the table", rpc_.tenant_denials());
    if (const auto tu = ks->tenant_usage(); !tu.empty()) {
      r.body += "# HELP bb_tenant_used_bytes bytes a tenant holds (size x replicas of its live objects)\n# TYPE bb_tenant_used_bytes gauge\n";
      for (const auto& u : tu) r.body += "bb_tenant_used_bytes{tenant=\"" + u.name + "\"} " + std::to_string(u.used_bytes) + "\n";
      r.body += "# HELP bb_tenant_objects live objects put by a tenant\n# TYPE bb_tenant_objects gauge\n";
      for (const auto& u : tu) r.body += "bb_tenant_objects{tenant=\"" + u.name + "\"} " + std::to_string(u.objects) + "\n";
      r.body += "# HELP bb_tenant_quota_bytes a tenant's byte budget (0 = unlimited)\n# TYPE bb_tenant_quota_bytes gauge\n";
      for (const auto& u : tu) r.body += "bb_tenant_quota_bytes{tenant=\"" + u.name + "\"} " + std::to_string(u.quota_bytes) + "\n";
    }
    r.body += "# TYPE bb_rpc_shm_channels gauge\nbb_rpc_shm_channels " + std::to_string(rpc_.shm_channels()) + "\n";
    return r;
  });
  http_.route("/healthz", [ks](const std::string&, const std::string&) {
    net::HttpResponse r;
    r.status = ks->is_running() ? 200 : 503;
    r.body = ks->is_running() ? (ks->is_leader() ? "ok leader\n" : "ok standby\n") : "stopped\n";
    return r;
  });
  http_.route("/stats", [ks](const std::string&, const std::string&) {
    net::HttpResponse r;
    r.content_type = "application/json";
    r.body = ks->stats_json().dump(2) + "\n";
    return r;
  });
}

ErrorCode RpcService::start() {
  if (running_) return ErrorCode::INVALID_STATE;
  auto hp = split_host_port(config_.listen_address);
  if (!hp) return ErrorCode::INVALID_ADDRESS;
  if (config_.rpc_busy_poll_us > 0) rpc_.set_busy_poll_us(config_.rpc_busy_poll_us);
  if (!config_.auth_token.empty()) net::set_cluster_token(config_.auth_token);
  if (config_.encrypt_transport) net::set_transport_encryption(true);
  if (!config_.auth_token_ro.empty()) net::set_cluster_token_ro(config_.auth_token_ro);
  if (!config_.http_auth_token.empty()) net::set_http_token(config_.http_auth_token);
  if (!config_.audit_log.empty() && !audit::open(config_.audit_log)) {
    BB_LOG(ERROR) << "keystone: cannot open the audit log " << config_.audit_log;
    return ErrorCode::INVALID_CONFIGURATION;
  }
  if (!config_.tenants_file.empty()) {
    std::string err;
    if (load_tenants_file(config_.tenants_file, &err) != ErrorCode::OK) {
      BB_LOG(ERROR) << "keystone: " << err;
      return ErrorCode::INVALID_CONFIGURATION;
    }
  } else {
    reload_tenants_if_changed();  // BB_TENANTS_FILE
  }
  if (!tenant_names().empty() && net::cluster_token().empty())
    BB_LOG(WARNING) << "keystone: tenants are configured but there is no cluster token -- connections that present nothing are still members";
  ErrorCode ec = rpc_.start(hp->first, static_cast<uint16_t>(hp->second), std::max(1, config_.rpc_threads));
  if (ec != ErrorCode::OK) return ec;
  if (!config_.http_metrics_port.empty() && config_.http_metrics_port != "off") {
    const uint16_t port = static_cast<uint16_t>(std::atoi(config_.http_metrics_port.c_str()));
    ec = http_.start(hp->first, port, 1);
    if (ec != ErrorCode::OK) {
      BB_LOG(WARNING) << "keystone: metrics port " << config_.http_metrics_port << " unavailable; continuing without HTTP";
    }
  }
  running_ = true;
  return ErrorCode::OK;
}

void RpcService::stop() {
  if (!running_) return;
  rpc_.stop();
  http_.stop();
  running_ = false;
}

// ================================================================ client
ErrorCode KeystoneRpcClient::connect(const std::string& host, uint16_t port, int timeout_ms) {
  return connect_any({host + ":" + std::to_string(port)}, timeout_ms);
}
ErrorCode KeystoneRpcClient::connect(const std::string& host_port, int timeout_ms) {
  std::vector<std::string> eps;
  size_t pos = 0;
  while (pos <= host_port.size()) {
    const size_t comma = std::min(host_port.find(',', pos), host_port.size());
    if (comma > pos) eps.push_back(host_port.substr(pos, comma - pos));
    pos = comma + 1;
  }
  return connect_any(eps, timeout_ms);
}
ErrorCode KeystoneRpcClient::connect_any(const std::vector<std::string>& endpoints, int timeout_ms) {
  std::vector<std::pair<std::string, uint16_t>> parsed;
  for (const auto& e : endpoints) {
    auto hp = split_host_port(e);
    if (!hp) return ErrorCode::INVALID_ADDRESS;
    parsed.emplace_back(hp->first, static_cast<uint16_t>(hp->second));
  }
  if (parsed.empty()) return ErrorCode::INVALID_ADDRESS;
  std::lock_guard<std::mutex> lk(ep_mu_);
  endpoints_ = std::move(parsed);
  connect_timeout_ms_ = timeout_ms;
  ErrorCode last = ErrorCode::CONNECTION_FAILED;
  for (size_t i = 0; i < endpoints_.size(); ++i) {
    auto c = std::make_shared<net::RpcClient>();
    last = c->connect(endpoints_[i].first, endpoints_[i].second, timeout_ms);
    if (last == ErrorCode::OK) {
      rpc_ = std::move(c);
      active_ = i;
      ++gen_;
      return ErrorCode::OK;
    }
  }
  return last;
}

bool KeystoneRpcClient::connected() const {
  std::lock_guard<std::mutex> lk(ep_mu_);
  return rpc_ && rpc_->connected();
}
std::string KeystoneRpcClient::active_endpoint() const {
  std::lock_guard<std::mutex> lk(ep_mu_);
  if (endpoints_.empty()) return {};
  return endpoints_[active_].first + ":" + std::to_string(endpoints_[active_].second);
}
std::shared_ptr<net::RpcClient> KeystoneRpcClient::current(uint64_t* gen) const {
  std::lock_guard<std::mutex> lk(ep_mu_);
  *gen = gen_;
  return rpc_;
}
void KeystoneRpcClient::rotate(uint64_t seen_gen) {
  std::lock_guard<std::mutex> lk(ep_mu_);
  if (gen_ != seen_gen) return;
  const size_t n = endpoints_.size();
  for (size_t step = 1; step <= n; ++step) {
    const size_t i = (active_ + step) % n;
    auto c = std::make_shared<net::RpcClient>();
    if (c->connect(endpoints_[i].first, endpoints_[i].second, connect_timeout_ms_) != ErrorCode::OK) continue;
    rpc_ = std::move(c);  // calls still in flight on the old connection keep their shared_ptr
    active_ = i;
    ++gen_;
    failovers_.fetch_add(1);
    return;
  }
}

Result<std::string> KeystoneRpcClient::call(uint32_t method, const std::string& req) {
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(failover_budget_ms_);
  for (size_t attempt = 1;; ++attempt) {
    uint64_t gen = 0;
    auto c = current(&gen);
    if (!c) return ErrorCode::CLIENT_DISCONNECTED;
    auto r = c->call(method, req, timeout_ms_);
    size_t n_eps = 0;
    {
      std::lock_guard<std::mutex> lk(ep_mu_);
      n_eps = endpoints_.size();
    }
    bool not_leader = false;
    if (r.ok() && r.value().size() >= 4) {
      Reader rd(r.value());
      not_leader = rd.ec() == ErrorCode::NOT_LEADER;
    }
    const bool retry = n_eps > 1 && (!r.ok() || not_leader) && std::chrono::steady_clock::now() < deadline;
    if (!retry) {
      if (!r.ok()) return r.error() == ErrorCode::CLIENT_DISCONNECTED ? ErrorCode::CLIENT_DISCONNECTED : ErrorCode::RPC_FAILED;
      return r;
    }
    rotate(gen);
    uint64_t now_gen = 0;
    current(&now_gen);
    // Nothing else reachable, or every keystone has been asked and nobody is leader yet (election in progress):
    // pace the next round.
    if (now_gen == gen || attempt % n_eps == 0) std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
}

#define BB_RPC(method, writer)                 \
  auto _resp = call(method, (writer).data());  \
  if (!_resp.ok()) return _resp.error();       \
  Reader rd(_resp.value());

Result<bool> KeystoneRpcClient::object_exists(const ObjectKey& key) {
  Writer w;
  w.str(key);
  BB_RPC(M_OBJECT_EXISTS, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return rd.boolean();
}
Result<std::vector<CopyPlacement>> KeystoneRpcClient::get_workers(const ObjectKey& key) {
  Writer w;
  w.str(key);
  BB_RPC(M_GET_WORKERS, w);
  return get_copies_result(rd);
}
Result<std::vector<CopyPlacement>> KeystoneRpcClient::put_start(const ObjectKey& key, size_t size, const WorkerConfig& cfg) {
  Writer w;
  w.str(key);
  w.u64(size);
  wire::put(w, cfg);
  w.str(client_id_);
  w.str(node_id_);
  BB_RPC(M_PUT_START, w);
  return get_copies_result(rd);
}
ErrorCode KeystoneRpcClient::put_complete(const ObjectKey& key, const ShardChecksums& sums) {
  Writer w;
  w.str(key);
  put_sums(w, sums);
  BB_RPC(M_PUT_COMPLETE, w);
  return rd.ec();
}
ErrorCode KeystoneRpcClient::put_cancel(const ObjectKey& key) {
  Writer w;
  w.str(key);
  BB_RPC(M_PUT_CANCEL, w);
  return rd.ec();
}
ErrorCode KeystoneRpcClient::remove_object(const ObjectKey& key) {
  Writer w;
  w.str(key);
  BB_RPC(M_REMOVE_OBJECT, w);
  return rd.ec();
}
ErrorCode KeystoneRpcClient::migrate_object(const ObjectKey& key, StorageClass target) {
  Writer w;
  w.str(key);
  w.u32(static_cast<uint32_t>(target));
  BB_RPC(M_MIGRATE_OBJECT, w);
  return rd.ec();
}
Result<std::vector<KeystoneApi::WorkerSummary>> KeystoneRpcClient::get_workers_info() {
  Writer w;
  BB_RPC(M_GET_WORKERS_INFO, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  std::vector<WorkerSummary> v(rd.count(16));
  for (auto& ws : v) {
    ws.worker_id = rd.str();
    ws.node_id = rd.str();
    ws.endpoint = rd.str();
    ws.heartbeat_age_ms = rd.i64();
    const uint32_t np = rd.count(4);
    for (uint32_t i = 0; i < np; ++i) ws.pools.push_back(rd.str());
  }
  if (!rd.ok()) return ErrorCode::RPC_FAILED;
  return v;
}
Result<std::vector<keystone::KeystoneService::ListedObject>> KeystoneRpcClient::list_objects(const std::string& prefix, size_t limit,
                                                                                           const std::string& start_after) {
  Writer w;
  w.str(prefix);
  w.u64(limit);
  w.str(start_after);
  BB_RPC(M_LIST_OBJECTS, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  std::vector<keystone::KeystoneService::ListedObject> v(rd.count(20));
  for (auto& o : v) {
    o.key = rd.str();
    o.size = rd.u64();
    o.copies = rd.u32();
    o.tier = static_cast<StorageClass>(rd.u32());
  }
  if (!rd.ok()) return ErrorCode::RPC_FAILED;
  return v;
}
Result<size_t> KeystoneRpcClient::compact_pool(const MemoryPoolId& pool, size_t max_moves) {
  Writer w;
  w.str(pool);
  w.u64(max_moves);
  BB_RPC(M_COMPACT_POOL, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return static_cast<size_t>(rd.u64());
}
Result<size_t> KeystoneRpcClient::drain_worker(const WorkerId& id) {
  Writer w;
  w.str(id);
  BB_RPC(M_DRAIN_WORKER, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return static_cast<size_t>(rd.u64());
}
Result<keystone::ScrubReport> KeystoneRpcClient::scrub(const std::string& prefix, size_t max_objects) {
  Writer w;
  w.str(prefix);
  w.u64(max_objects);
  BB_RPC(M_SCRUB, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  keystone::ScrubReport rep;
  for (uint64_t* v : {&rep.objects, &rep.copies, &rep.corrupt, &rep.healed, &rep.unrecoverable, &rep.unreachable}) *v = rd.u64();
  if (!rd.ok()) return ErrorCode::RPC_FAILED;
  return rep;
}
Result<std::vector<keystone::TenantUsage>> KeystoneRpcClient::tenant_usage() {
  Writer w;
  BB_RPC(M_TENANT_USAGE, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  std::vector<keystone::TenantUsage> v(rd.count(36));
  for (auto& u : v) {
    u.name = rd.str();
    for (uint64_t* x : {&u.used_bytes, &u.objects, &u.quota_bytes, &u.max_objects}) *x = rd.u64();
  }
  if (!rd.ok()) return ErrorCode::RPC_FAILED;
  return v;
}
ErrorCode KeystoneRpcClient::remove_worker(const WorkerId& id) {
  Writer w;
  w.str(id);
  BB_RPC(M_REMOVE_WORKER, w);
  return rd.ec();
}
Result<size_t> KeystoneRpcClient::remove_all_objects() {
  Writer w;
  BB_RPC(M_REMOVE_ALL_OBJECTS, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return static_cast<size_t>(rd.u64());
}
Result<ClusterStats> KeystoneRpcClient::get_cluster_stats() {
  Writer w;
  BB_RPC(M_GET_CLUSTER_STATS, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  ClusterStats s;
  wire::get(rd, s);
  return s;
}
Result<ViewVersionId> KeystoneRpcClient::get_view_version() {
  Writer w;
  BB_RPC(M_GET_VIEW_VERSION, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return static_cast<ViewVersionId>(rd.i64());
}

std::vector<Result<bool>> KeystoneRpcClient::batch_object_exists(const std::vector<ObjectKey>& keys) {
  Writer w;
  put_keys(w, keys);
  auto resp = call(M_BATCH_OBJECT_EXISTS, w.data());
  std::vector<Result<bool>> out;
  if (!resp.ok()) {
    out.assign(keys.size(), Result<bool>(resp.error()));
    return out;
  }
  Reader rd(resp.value());
  if (const ErrorCode ec = rd.ec(); ec != ErrorCode::OK) {  // e.g. NOT_LEADER for the whole batch
    out.assign(keys.size(), Result<bool>(ec));
    return out;
  }
  const uint32_t n = rd.count(5);
  for (uint32_t i = 0; i < n; ++i) {
    const ErrorCode ec = rd.ec();
    const bool b = rd.boolean();
    out.push_back(ec == ErrorCode::OK ? Result<bool>(b) : Result<bool>(ec));
  }
  out.resize(keys.size(), Result<bool>(ErrorCode::RPC_FAILED));
  return out;
}

std::vector<Result<std::vector<CopyPlacement>>> KeystoneRpcClient::batch_get_workers(const std::vector<ObjectKey>& keys) {
  Writer w;
  put_keys(w, keys);
  auto resp = call(M_BATCH_GET_WORKERS, w.data());
  std::vector<Result<std::vector<CopyPlacement>>> out;
  if (!resp.ok()) {
    out.assign(keys.size(), Result<std::vector<CopyPlacement>>(resp.error()));
    return out;
  }
  Reader rd(resp.value());
  if (const ErrorCode ec = rd.ec(); ec != ErrorCode::OK) {
    out.assign(keys.size(), Result<std::vector<CopyPlacement>>(ec));
    return out;
  }
  const uint32_t n = rd.count(4);
  out.reserve(keys.size());
  wire::PlacementBatchReader pr(rd);
  for (uint32_t i = 0; i < n; ++i) out.push_back(pr.get());
  out.resize(keys.size(), Result<std::vector<CopyPlacement>>(ErrorCode::RPC_FAILED));
  return out;
}

std::vector<Result<std::vector<CopyPlacement>>> KeystoneRpcClient::batch_put_start(const std::vector<PutStartItem>& items) {
  Writer w;
  w.u32(static_cast<uint32_t>(items.size()));
  bool one_config = !items.empty();
  for (size_t i = 1; i < items.size() && one_config; ++i) one_config = items[i].config == items[0].config;
  w.boolean(one_config);
  if (one_config) wire::put(w, items[0].config);
  for (const auto& it : items) {
    w.str(it.key);
    w.u64(it.size);
    if (!one_config) wire::put(w, it.config);
  }
  w.str(client_id_);
  w.str(node_id_);
  auto resp = call(M_BATCH_PUT_START, w.data());
  std::vector<Result<std::vector<CopyPlacement>>> out;
  if (!resp.ok()) {
    out.assign(items.size(), Result<std::vector<CopyPlacement>>(resp.error()));
    return out;
  }
  Reader rd(resp.value());
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) {
    out.assign(items.size(), Result<std::vector<CopyPlacement>>(ec));
    return out;
  }
  const uint32_t n = rd.count(4);
  out.reserve(items.size());
  wire::PlacementBatchReader pr(rd);
  for (uint32_t i = 0; i < n; ++i) out.push_back(pr.get());
  out.resize(items.size(), Result<std::vector<CopyPlacement>>(ErrorCode::RPC_FAILED));
  return out;
}

namespace {
std::vector<ErrorCode> parse_ecs(const Result<std::string>& resp, size_t n_expected) {
  std::vector<ErrorCode> out;
  if (!resp.ok()) {
    out.assign(n_expected, resp.error());
    return out;
  }
  Reader rd(resp.value());
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) {
    out.assign(n_expected, ec);
    return out;
  }
  const uint32_t n = rd.count(4);
  for (uint32_t i = 0; i < n; ++i) out.push_back(rd.ec());
  out.resize(n_expected, ErrorCode::RPC_FAILED);
  return out;
}
}  // namespace

std::vector<ErrorCode> KeystoneRpcClient::batch_put_complete(const std::vector<ObjectKey>& keys, const std::vector<ShardChecksums>& sums) {
  Writer w;
  put_keys(w, keys);
  w.u32(static_cast<uint32_t>(sums.size()));
  for (const auto& s : sums) put_sums(w, s);
  return parse_ecs(call(M_BATCH_PUT_COMPLETE, w.data()), keys.size());
}
std::vector<ErrorCode> KeystoneRpcClient::batch_put_cancel(const std::vector<ObjectKey>& keys) {
  Writer w;
  put_keys(w, keys);
  return parse_ecs(call(M_BATCH_PUT_CANCEL, w.data()), keys.size());
}
std::vector<ErrorCode> KeystoneRpcClient::batch_remove_object(const std::vector<ObjectKey>& keys) {
  Writer w;
  put_keys(w, keys);
  return parse_ecs(call(M_BATCH_REMOVE_OBJECT, w.data()), keys.size());
}

Result<std::vector<MemoryPool>> KeystoneRpcClient::get_memory_pools() {
  Writer w;
  BB_RPC(M_GET_MEMORY_POOLS, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  std::vector<MemoryPool> v(rd.count(16));
  for (auto& p : v) wire::get(rd, p);
  if (!rd.ok()) return ErrorCode::RPC_FAILED;
  return v;
}
Result<std::string> KeystoneRpcClient::client_register(const std::string& node_id) {
  Writer w;
  w.str(node_id);
  BB_RPC(M_CLIENT_REGISTER, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return rd.str();
}
Result<ViewVersionId> KeystoneRpcClient::client_ping(const std::string& client_id) {
  Writer w;
  w.str(client_id);
  BB_RPC(M_CLIENT_PING, w);
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return static_cast<ViewVersionId>(rd.i64());
}
ErrorCode KeystoneRpcClient::register_worker(const WorkerRecord& rec) {
  Writer w;
  w.str(to_json(rec).dump());
  BB_RPC(M_REGISTER_WORKER, w);
  return rd.ec();
}
ErrorCode KeystoneRpcClient::register_memory_pool(const MemoryPool& pool) {
  Writer w;
  wire::put(w, pool);
  BB_RPC(M_REGISTER_MEMORY_POOL, w);
  return rd.ec();
}
ErrorCode KeystoneRpcClient::worker_heartbeat(const WorkerId& id) {
  Writer w;
  w.str(id);
  BB_RPC(M_WORKER_HEARTBEAT, w);
  return rd.ec();
}

// ================================================================ bootstrap
Result<KeystoneBundle> create_and_start_keystone(const KeystoneConfig& config) {
  KeystoneBundle b;
  std::string err;
  ErrorCode ec = config.validate(&err);
  if (ec != ErrorCode::OK) {
    BB_LOG(ERROR) << "keystone config invalid: " << err;
    return ec;
  }
  if (!config.etcd_endpoints.empty() && config.etcd_endpoints != "none") {
    b.coord = std::make_shared<coord::CoordService>(config.etcd_endpoints);
    ec = b.coord->connect();
    if (ec != ErrorCode::OK) {
      BB_LOG(ERROR) << "cannot connect to coordination endpoints " << config.etcd_endpoints;
      return ec;
    }
  }
  b.keystone = std::make_shared<keystone::KeystoneService>(config, b.coord);
  ec = b.keystone->initialize();
  if (ec != ErrorCode::OK) return ec;
  ec = b.keystone->start();
  if (ec != ErrorCode::OK) return ec;
  b.keystone->set_copy_verifier(client::make_data_server_verifier());  // scrub hashes copies at their workers
  b.keystone->set_copy_mover(client::make_data_server_mover());  // tier demotion + re-replication move real bytes
  b.keystone->set_reservation_hooks(client::make_data_server_reservation_hooks());  // used when enable_reservations is set
  b.rpc = std::make_unique<RpcService>(b.keystone, config);
  ec = b.rpc->start();
  if (ec != ErrorCode::OK) {
    b.keystone->stop();
    return ec;
  }
  return b;
}

}  // namespace bb::rpc
