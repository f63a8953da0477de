#include "common/fault.h"
#include "keystone/keystone_service.h"

#include "common/audit.h"

#include <unordered_set>

#include <algorithm>
#include <chrono>
#include <map>
#include <set>

#include "common/log.h"
#include "rpc/wire.h"

namespace bb::keystone {

namespace {
int64_t wall_ms() {
  return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}
double us_since(TimePoint t0) {
  return std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
}
constexpr auto kPendingGrace = std::chrono::minutes(10);

std::string encode_object(const ObjectInfo& o) {
  wire::Writer w;
  w.str(o.key);
  w.u64(o.size);
  const int64_t age_ms = std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now() - o.created).count();
  w.i64(wall_ms() - age_ms);  // wall-clock creation time survives a leader change
  wire::put(w, o.config);
  wire::put(w, o.copies);
  w.u32(static_cast<uint32_t>(o.state));
  w.str(o.tenant);  // (records written before tenants existed end here)
  return w.take();
}
bool decode_object(const std::string& s, ObjectInfo& o) {
  wire::Reader r(s);
  o.key = r.str();
  o.size = r.u64();
  const int64_t created_wall = r.i64();
  wire::get(r, o.config);
  wire::get(r, o.copies);
  o.state = static_cast<ObjectState>(r.u32());
  if (!r.ok()) return false;
  if (!r.at_end()) o.tenant = r.str();
  if (!r.ok()) return false;
  const int64_t age = std::max<int64_t>(0, wall_ms() - created_wall);
  o.created = Clock::now() - std::chrono::milliseconds(age);
  o.last_accessed = Clock::now();
  return true;
}
}  // namespace

KeystoneService::KeystoneService(const KeystoneConfig& config, std::shared_ptr<coord::CoordService> coord)
    : config_(config), coord_(std::move(coord)),
      allocator_(std::make_unique<alloc::KeystoneAllocatorAdapter>(alloc::AllocatorFactory::create_range_based())) {
  metrics_.describe("put_start_total", "put_start calls that allocated placements");
  metrics_.describe("put_complete_total", "objects that reached COMPLETE");
  metrics_.describe("get_workers_total", "successful placement lookups");
  metrics_.describe("evictions_total", "objects dropped by watermark eviction");
  metrics_.describe("demotions_total", "objects moved to a lower tier by watermark eviction");
  metrics_.describe("expired_total", "objects reclaimed by TTL");
  metrics_.describe("worker_deaths_total", "workers removed after heartbeat lease expiry");
  metrics_.describe("put_bytes_total", "logical bytes admitted by put_start");
  // series touched on every object: resolve them once (Metrics::inc looks the name up under a lock)
  hot_.put_start_total = metrics_.counter_ref("put_start_total");
  hot_.put_start_failed_total = metrics_.counter_ref("put_start_failed_total");
  hot_.put_bytes_total = metrics_.counter_ref("put_bytes_total");
  hot_.put_complete_total = metrics_.counter_ref("put_complete_total");
  hot_.get_workers_total = metrics_.counter_ref("get_workers_total");
  hot_.remove_total = metrics_.counter_ref("remove_total");
  hot_.put_start_latency = metrics_.histogram_ref("put_start_latency_us");
}

KeystoneService::~KeystoneService() { stop(); }

// ================================================================ lifecycle
ErrorCode KeystoneService::initialize() {
  if (config_.service_id.empty()) config_.service_id = "keystone-" + uuid_to_string(generate_uuid()).substr(0, 12);
  candidate_id_ = config_.service_id;
  if (coord_) {
    ErrorCode ec = setup_coordination();
    if (ec != ErrorCode::OK) return ec;
  }
  return ErrorCode::OK;
}

ErrorCode KeystoneService::setup_coordination() {
  ErrorCode ec = coord_->connect();
  if (ec != ErrorCode::OK) {
    BB_LOG(ERROR) << "keystone: cannot reach coordination store";
    return ec;
  }
  ec = coord_->register_service("blackbird-keystone", config_.service_id, config_.listen_address, config_.service_registration_ttl_sec);
  if (ec != ErrorCode::OK) return ec;
  if (config_.enable_ha) {
    bool won = false;
    const TimePoint t0 = Clock::now();
    ec = coord_->campaign_leader(election_name(), candidate_id_, config_.service_registration_ttl_sec, won);
    if (ec != ErrorCode::OK) return ec;
    if (won) {
      term_.store(coord_->leader_term(election_name()));
      arm_lease_deadline(t0);
    }
    leader_.store(won, std::memory_order_release);
    BB_LOG(INFO) << "keystone " << candidate_id_ << (won ? " is the leader (term " + std::to_string(term_.load()) + ")" : " is a standby");
  }
  return ErrorCode::OK;
}

// Leadership is valid in local time until `refreshed_at` (taken BEFORE the keep-alive round trip) + TTL - margin.
void KeystoneService::arm_lease_deadline(TimePoint refreshed_at) {
  const auto ttl = std::chrono::seconds(std::max<int64_t>(1, config_.service_registration_ttl_sec));
  const auto margin = std::max<std::chrono::nanoseconds>(std::chrono::milliseconds(200), ttl / 10);
  lease_deadline_.store((refreshed_at + ttl - margin).time_since_epoch().count(), std::memory_order_release);
}

void KeystoneService::step_down(const char* why) {
  if (!leader_.exchange(false)) return;
  BB_LOG(WARNING) << "keystone " << candidate_id_ << " steps down: " << why;
  term_.store(0);
  lease_deadline_.store(0);
  metrics_.inc("leadership_lost_total");
}

// A non-leader holds no authoritative state: whatever it remembers may be changed by the next leader, so the table and
// the allocator are rebuilt from the metadata log when (if) leadership comes back -- never patched up.
void KeystoneService::reset_object_state() {
  for (auto& sh : shards_) {
    std::lock_guard<std::mutex> wl(sh.wal_mu);
    std::lock_guard<SpinMutex> lk(sh.mu);
    sh.objects.clear();
    for (auto& op : sh.wal_queue)
      if (op.result) *op.result = ErrorCode::NOT_LEADER;
    sh.wal_queue.clear();
  }
  static_cast<alloc::RangeAllocator&>(allocator_->allocator()).reset();  // in place: other threads hold the adapter
  {
    std::lock_guard<std::mutex> tl(tenant_mu_);
    tenant_usage_.clear();
  }
  std::lock_guard<std::mutex> ul(unadopted_mu_);
  unadopted_.clear();
}

ErrorCode KeystoneService::start() {
  if (running_.exchange(true)) return ErrorCode::INVALID_STATE;
  view_version_.store(1);
  if (coord_ && coord_->is_connected()) {
    load_existing_state();
    const std::string p = cluster_prefix();
    int64_t id = 0;
    if (coord_->watch_prefix(p + "workers/", [this](const std::string& k, const std::string& v, bool d) { on_worker_event(k, v, d); }, &id) == ErrorCode::OK)
      watch_ids_.push_back(id);
    if (coord_->watch_prefix(p + "memory_pools/", [this](const std::string& k, const std::string& v, bool d) { on_legacy_pool_event(k, v, d); }, &id) == ErrorCode::OK)
      watch_ids_.push_back(id);
    if (coord_->watch_prefix(p + "heartbeat/", [this](const std::string& k, const std::string& v, bool d) { on_heartbeat_event(k, v, d); }, &id) == ErrorCode::OK)
      watch_ids_.push_back(id);
    if (config_.enable_reservations &&
        coord_->watch_prefix(p + "reservations_expired/", [this](const std::string& k, const std::string& v, bool d) { on_reservation_expired(k, v, d); }, &id) == ErrorCode::OK)
      watch_ids_.push_back(id);
  }
  if (config_.enable_ha) {
    wal_on_.store(coord_ && coord_->is_connected());
    if (is_leader()) recover_objects_from_wal();
  } else if (!config_.wal_path.empty()) {
    const ErrorCode ec = open_local_wal();  // replays the local log into the table
    if (ec != ErrorCode::OK) {
      running_.store(false);
      return ec;
    }
    wal_on_.store(true);
  }
  if (config_.enable_gc) gc_thread_ = std::thread([this] { gc_loop(); });
  health_thread_ = std::thread([this] { health_loop(); });
  if (coord_) keepalive_thread_ = std::thread([this] { keepalive_loop(); });
  return ErrorCode::OK;
}

void KeystoneService::stop() {
  if (!running_.exchange(false)) return;
  {
    std::lock_guard<std::mutex> lk(sleep_mu_);
    sleep_cv_.notify_all();
  }
  if (gc_thread_.joinable()) gc_thread_.join();
  if (health_thread_.joinable()) health_thread_.join();
  if (keepalive_thread_.joinable()) keepalive_thread_.join();
  // The watch callbacks capture `this`: unwatch() returns only once no callback is running and none will start, so the
  // coordination client's push thread can never call into a stopped (possibly destroyed) service.
  if (coord_)
    for (int64_t id : watch_ids_) coord_->unwatch(id);
  watch_ids_.clear();
  if (coord_ && coord_->is_connected()) {
    if (config_.enable_ha && leader_.load()) coord_->resign_leader("keystone-" + config_.cluster_id, candidate_id_);
    coord_->unregister_service("blackbird-keystone", config_.service_id);
  }
  leader_.store(false);
  if (local_wal_) {
    maybe_snapshot_local_wal();
    local_wal_->close();
  }
}

bool KeystoneService::interruptible_sleep(std::chrono::milliseconds d) {
  std::unique_lock<std::mutex> lk(sleep_mu_);
  sleep_cv_.wait_for(lk, d, [this] { return !running_.load(); });
  return running_.load();
}

void KeystoneService::gc_loop() {
  while (interruptible_sleep(std::chrono::seconds(config_.gc_interval_sec))) {
    if (is_leader()) run_gc_once();
  }
}

void KeystoneService::health_loop() {
  while (interruptible_sleep(std::chrono::seconds(config_.health_check_interval_sec))) {
    if (!is_leader()) continue;
    // workers registered without a coordination lease are aged out by heartbeat timestamps
    std::vector<WorkerId> stale;
    if (!coord_) {
      std::shared_lock<std::shared_mutex> lk(workers_mu_);
      for (const auto& [id, w] : workers_)
        if (w.is_stale(std::chrono::seconds(config_.worker_heartbeat_ttl_sec))) stale.push_back(id);
    }
    for (const auto& id : stale) handle_worker_death(id);
    // expire client sessions
    {
      std::lock_guard<std::mutex> lk(clients_mu_);
      const auto now = Clock::now();
      for (auto it = clients_.begin(); it != clients_.end();)
        it = (now - it->second.last_ping > std::chrono::seconds(config_.client_ttl_sec)) ? clients_.erase(it) : std::next(it);
    }
    if (reload_tenants_if_changed())  // an edited tenant table (grants, budgets, revocations) takes effect without a restart
      audit::event("tenants_reloaded", {{"who", "keystone"}, {"count", std::to_string(tenant_names().size())}});
    run_eviction_once();
    run_repair_once();
    run_promotion_once();
    run_compaction_once();
    if (config_.scrub_objects_per_round > 0) scrub_from("", static_cast<size_t>(config_.scrub_objects_per_round), scrub_cursor_);
  }
}

void KeystoneService::keepalive_loop() {
  const std::string election = election_name();
  // HA: the election lease must be refreshed several times per TTL, whatever the (registry) refresh interval says
  const int64_t ttl_ms = std::max<int64_t>(1, config_.service_registration_ttl_sec) * 1000;
  const int64_t period_ms = config_.enable_ha ? std::min<int64_t>(config_.service_refresh_interval_sec * 1000, std::max<int64_t>(100, ttl_ms / 4))
                                              : config_.service_refresh_interval_sec * 1000;
  bool was_leader = leader_.load();
  int backoff = 0;
  while (interruptible_sleep(std::chrono::milliseconds(period_ms))) {
    if (!coord_ || !coord_->is_connected()) continue;
    if (coord_->register_service("blackbird-keystone", config_.service_id, config_.listen_address,
                                 config_.service_registration_ttl_sec) != ErrorCode::OK)
      BB_LOG(WARNING) << "keystone: service registration refresh failed";
    if (!config_.enable_ha) continue;
    if (leader_.load()) {
      const TimePoint t0 = Clock::now();
      const ErrorCode ec = coord_->refresh_leadership(election, candidate_id_);
      if (ec == ErrorCode::OK) arm_lease_deadline(t0);
      else if (ec != ErrorCode::ETCD_ERROR) step_down("the election lease is gone");
      // ETCD_ERROR: the store is unreachable -- leadership is unknown; is_leader() lapses by itself at the deadline
    }
    if (was_leader && !leader_.load()) {  // also reached when a fenced write stepped us down
      reset_object_state();
      was_leader = false;
      backoff = 2;  // a deposed leader sits out two rounds: whoever replaced it (or a healthy standby) goes first
      continue;
    }
    was_leader = leader_.load();
    if (backoff > 0) {
      --backoff;
      continue;
    }
    if (!leader_.load()) {
      bool won = false;
      const TimePoint t0 = Clock::now();
      if (coord_->campaign_leader(election, candidate_id_, config_.service_registration_ttl_sec, won) == ErrorCode::OK && won) {
        term_.store(coord_->leader_term(election));
        BB_LOG(INFO) << "keystone " << candidate_id_ << " took over leadership (term " << term_.load() << ")";
        reset_object_state();
        load_existing_state();
        recover_objects_from_wal();
        arm_lease_deadline(t0);
        leader_.store(true, std::memory_order_release);
        was_leader = true;
        bump_view();
      }
    }
  }
}

// ================================================================ coordination events
void KeystoneService::load_existing_state() {
  std::vector<std::string> keys, values;
  const std::string p = cluster_prefix();
  if (coord_->get_with_prefix(p + "workers/", keys, values) == ErrorCode::OK)
    for (size_t i = 0; i < keys.size(); ++i) on_worker_event(keys[i], values[i], false);
  if (coord_->get_with_prefix(p + "memory_pools/", keys, values) == ErrorCode::OK)
    for (size_t i = 0; i < keys.size(); ++i) on_legacy_pool_event(keys[i], values[i], false);
}

void KeystoneService::on_worker_event(const std::string& key, const std::string& value, bool is_delete) {
  const std::string base = cluster_prefix() + "workers/";
  if (key.compare(0, base.size(), base) != 0) return;
  const std::string rest = key.substr(base.size());
  const size_t mp = rest.find("/memory_pools/");
  if (mp != std::string::npos) {
    const std::string pid = rest.substr(mp + 14);
    if (is_delete) {
      handle_pool_removed(pid);
      return;
    }
    auto j = Json::parse(value);
    if (!j) return;
    auto pool = memory_pool_from_json(*j);
    if (!pool.ok()) {
      BB_LOG(WARNING) << "keystone: malformed pool record at " << key;
      return;
    }
    if (pool.value().worker_id.empty()) pool.value().worker_id = rest.substr(0, mp);
    register_memory_pool(pool.value());
    return;
  }
  if (rest.find('/') != std::string::npos) return;
  if (is_delete) {
    // explicit deregistration (clean shutdown): same handling as death, minus the metric
    handle_worker_death(rest);
    return;
  }
  auto j = Json::parse(value);
  if (!j) return;
  auto rec = worker_record_from_json(*j);
  if (!rec.ok()) return;
  if (rec.value().worker_id.empty()) rec.value().worker_id = rest;
  register_worker(rec.value());
}

void KeystoneService::on_legacy_pool_event(const std::string& key, const std::string& value, bool is_delete) {
  const std::string base = cluster_prefix() + "memory_pools/";
  if (key.compare(0, base.size(), base) != 0) return;
  const std::string pid = key.substr(base.size());
  if (is_delete) {
    handle_pool_removed(pid);
    return;
  }
  auto j = Json::parse(value);
  if (!j) return;
  auto pool = memory_pool_from_json(*j);
  if (pool.ok()) register_memory_pool(pool.value());
}

void KeystoneService::on_heartbeat_event(const std::string& key, const std::string&, bool is_delete) {
  const std::string base = cluster_prefix() + "heartbeat/";
  if (key.compare(0, base.size(), base) != 0) return;
  const std::string wid = key.substr(base.size());
  if (!is_delete) {
    worker_heartbeat(wid);
    return;
  }
  {
    std::shared_lock<std::shared_mutex> lk(workers_mu_);
    if (!workers_.count(wid)) return;  // already deregistered (clean shutdown deletes the worker key first)
  }
  BB_LOG(WARNING) << "keystone: heartbeat lease of worker " << wid << " expired";
  metrics_.inc("worker_deaths_total");
  handle_worker_death(wid);
  if (coord_ && is_leader()) {
    const std::string p = cluster_prefix() + "workers/" + wid;
    if (auto st = coord_->store()) st->del_prefix(p + "/");
    coord_->del(p);
  }
}

// ================================================================ registries
ErrorCode KeystoneService::register_worker(const WorkerRecord& rec) {
  if (rec.worker_id.empty()) return ErrorCode::INVALID_WORKER;
  {
    std::unique_lock<std::shared_mutex> lk(workers_mu_);
    WorkerInfo& w = workers_[rec.worker_id];
    w.worker_id = rec.worker_id;
    w.node_id = rec.node_id;
    w.endpoint = rec.rpc_endpoint;
    w.record = rec;
    w.last_heartbeat = Clock::now();
  }
  bump_view();
  return ErrorCode::OK;
}

ErrorCode KeystoneService::register_memory_pool(const MemoryPool& pool) {
  if (pool.id.empty() || pool.size == 0) return ErrorCode::INVALID_MEMORY_POOL;
  {
    std::unique_lock<std::shared_mutex> lk(workers_mu_);
    if (!pool.worker_id.empty()) {
      WorkerInfo& w = workers_[pool.worker_id];
      if (w.worker_id.empty()) {
        w.worker_id = pool.worker_id;
        w.node_id = pool.node_id;
        w.last_heartbeat = Clock::now();
      }
      if (std::find(w.pools.begin(), w.pools.end(), pool.id) == w.pools.end()) w.pools.push_back(pool.id);
    }
    std::unique_lock<std::shared_mutex> pk(pools_mu_);
    pools_[pool.id] = pool;
    refresh_top_tier_locked();
  }
  // objects recovered from the metadata log before this pool was known: reserve their extents now
  std::vector<ObjectKey> waiting;
  {
    std::lock_guard<std::mutex> ul(unadopted_mu_);
    auto it = unadopted_.find(pool.id);
    if (it != unadopted_.end()) {
      waiting = std::move(it->second);
      unadopted_.erase(it);
    }
  }
  if (!waiting.empty()) {
    std::shared_lock<std::shared_mutex> pk(pools_mu_);
    auto& ra = static_cast<alloc::RangeAllocator&>(allocator_->allocator());
    for (const auto& key : waiting) {
      auto info = get_object_info(key);
      if (info.ok()) ra.adopt(key, info.value().copies, pools_, pool.id);
    }
  }
  bump_view();
  return ErrorCode::OK;
}

ErrorCode KeystoneService::worker_heartbeat(const WorkerId& id) {
  std::unique_lock<std::shared_mutex> lk(workers_mu_);
  auto it = workers_.find(id);
  if (it == workers_.end()) return ErrorCode::INVALID_WORKER;
  it->second.last_heartbeat = Clock::now();
  return ErrorCode::OK;
}

ErrorCode KeystoneService::remove_worker(const WorkerId& id) {
  {
    std::shared_lock<std::shared_mutex> lk(workers_mu_);
    if (!workers_.count(id)) return ErrorCode::INVALID_WORKER;
  }
  handle_worker_death(id);
  if (coord_ && coord_->is_connected()) {
    const std::string p = cluster_prefix();
    if (auto st = coord_->store()) st->del_prefix(p + "workers/" + id + "/");
    coord_->del(p + "workers/" + id);
    coord_->del(p + "heartbeat/" + id);
  }
  return ErrorCode::OK;
}

Result<size_t> KeystoneService::drain_worker(const WorkerId& id) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  if (!mover) return ErrorCode::NOT_IMPLEMENTED;
  std::vector<MemoryPoolId> mine;
  {
    std::shared_lock<std::shared_mutex> wl(workers_mu_);
    auto it = workers_.find(id);
    if (it == workers_.end()) return ErrorCode::INVALID_WORKER;
    mine = it->second.pools;
  }
  {
    std::unique_lock<std::shared_mutex> pk(pools_mu_);
    for (const auto& p : mine)
      if (std::find(draining_.begin(), draining_.end(), p) == draining_.end()) draining_.push_back(p);
  }
  bump_view();
  size_t moved = 0, stuck = 0;
  std::unordered_set<ObjectKey> handled;
  for (const auto& pid : mine) {
    {
      std::shared_lock<std::shared_mutex> pk(pools_mu_);
      if (!pools_.count(pid)) continue;
    }
    for (const ObjectKey& raw : allocator_->allocator().objects_on_pool(pid)) {
      const ObjectKey key = raw.substr(0, raw.find('\x01'));  // ledgers of moved copies are "<key>\x01<slot>"
      if (!handled.insert(key).second) continue;  // several shards / ledgers of one object: handled once, as a whole
      // Only the copies that touch this worker move (a replica elsewhere stays where it is, and keeps serving); one copy per
      // turn, metadata re-read after every swap.  The copy that leaves is its own best source -- same bytes, and on a GPU
      // worker the pull kernel -- and one that fails its digest on the way is rebuilt from a sibling instead.
      bool any = false;
      ErrorCode last = ErrorCode::OK;
      for (int turn = 0; turn < 64; ++turn) {
        auto info = get_object_info(key);
        if (!info.ok() || info.value().state != ObjectState::COMPLETE) {
          last = info.ok() ? ErrorCode::OBJECT_NOT_READY : info.error();
          break;
        }
        const auto& copies = info.value().copies;
        size_t idx = copies.size();
        for (size_t c = 0; c < copies.size() && idx == copies.size(); ++c)
          for (const auto& sp : copies[c].shards)
            if (std::find(mine.begin(), mine.end(), sp.pool_id) != mine.end()) {
              idx = c;
              break;
            }
        if (idx == copies.size()) break;  // nothing of it is left here
        last = replace_copy(mover, info.value(), idx, copies[idx]);
        for (size_t c = 0; last == ErrorCode::CHECKSUM_MISMATCH && c < copies.size(); ++c)
          if (c != idx) last = replace_copy(mover, info.value(), idx, copies[c]);
        if (last != ErrorCode::OK) break;
        any = true;
      }
      if (last == ErrorCode::OK) {
        if (any) {
          ++moved;
          metrics_.inc("drain_moves_total");
        }
      } else if (last != ErrorCode::OBJECT_NOT_FOUND && last != ErrorCode::OBJECT_NOT_READY) {
        ++stuck;
        BB_LOG(WARNING) << "drain " << id << ": " << key << " stays on " << pid << " (" << to_string(last) << ")";
      }
    }
  }
  if (moved) bump_view();
  if (stuck) return ErrorCode::INSUFFICIENT_SPACE;  // still draining: free space elsewhere and call again
  const ErrorCode rc = remove_worker(id);
  {
    std::unique_lock<std::shared_mutex> pk(pools_mu_);
    draining_.erase(std::remove_if(draining_.begin(), draining_.end(), [&](const MemoryPoolId& p) { return std::find(mine.begin(), mine.end(), p) != mine.end(); }),
                    draining_.end());
  }
  if (rc != ErrorCode::OK) return rc;
  return moved;
}

void KeystoneService::handle_pool_removed(const MemoryPoolId& pid) {
  {
    std::unique_lock<std::shared_mutex> lk(workers_mu_);
    std::unique_lock<std::shared_mutex> pk(pools_mu_);
    auto pit = pools_.find(pid);
    if (pit == pools_.end()) return;
    auto wit = workers_.find(pit->second.worker_id);
    if (wit != workers_.end()) {
      auto& v = wit->second.pools;
      v.erase(std::remove(v.begin(), v.end(), pid), v.end());
    }
    pools_.erase(pit);
    refresh_top_tier_locked();
  }
  invalidate_pools({pid}, "pool " + pid + " deregistered");
}

void KeystoneService::handle_worker_death(const WorkerId& id) {
  std::vector<MemoryPoolId> dead;
  {
    std::unique_lock<std::shared_mutex> lk(workers_mu_);
    auto it = workers_.find(id);
    if (it == workers_.end()) return;
    dead = it->second.pools;
    workers_.erase(it);
    std::unique_lock<std::shared_mutex> pk(pools_mu_);
    for (auto pit = pools_.begin(); pit != pools_.end();) {
      if (pit->second.worker_id == id && std::find(dead.begin(), dead.end(), pit->first) == dead.end()) dead.push_back(pit->first);
      ++pit;
    }
    for (const auto& p : dead) pools_.erase(p);
    refresh_top_tier_locked();
  }
  invalidate_pools(dead, "worker " + id + " died");
}

// The pools are gone from the registry: forget their allocators and invalidate every copy that had a shard on them.
void KeystoneService::invalidate_pools(const std::vector<MemoryPoolId>& dead, const std::string& why) {
  for (const auto& p : dead) allocator_->allocator().forget_pool(p);
  // invalidate every copy that had a shard on the dead pools
  const std::set<MemoryPoolId> dead_set(dead.begin(), dead.end());
  size_t lost_objects = 0, degraded = 0;
  for (auto& sh : shards_) {
    std::vector<ObjectKey> gone;
    ShardGuard lk(this, sh);
    for (auto& [key, info] : sh.objects) {
      const size_t before = info.copies.size();
      info.copies.erase(std::remove_if(info.copies.begin(), info.copies.end(),
                                       [&](const CopyPlacement& c) {
                                         for (const auto& s : c.shards)
                                           if (dead_set.count(s.pool_id)) return true;
                                         return false;
                                       }),
                        info.copies.end());
      if (info.copies.size() == before) continue;
      if (info.copies.empty()) gone.push_back(key);
      else {
        ++degraded;
        if (info.state == ObjectState::COMPLETE) persist_object(sh, info);
      }
    }
    for (const auto& k : gone) {
      erase_locked(sh, k, true);
      ++lost_objects;
    }
  }
  if (lost_objects || degraded) BB_LOG(WARNING) << why << ": " << lost_objects << " objects lost, " << degraded << " degraded";
  metrics_.inc("objects_lost_total", lost_objects);
  bump_view();
}

bool KeystoneService::pool_alive(const MemoryPoolId& id) const {
  std::shared_lock<std::shared_mutex> lk(pools_mu_);
  return pools_.count(id) > 0;
}

// ================================================================ object API
ErrorCode KeystoneService::erase_locked(Shard& sh, const ObjectKey& key, bool free_ranges) {
  auto it = sh.objects.find(key);
  if (it == sh.objects.end()) return ErrorCode::OBJECT_NOT_FOUND;
  const bool was_complete = it->second.state == ObjectState::COMPLETE || it->second.committing;
  const std::vector<std::string> extra = std::move(it->second.extra_ledgers);
  if (!it->second.tenant.empty()) tenant_release(it->second.tenant, tenant_charge_of(it->second));
  if (it->second.reserved && reservations_enabled()) {
    HookOp op;
    op.release = it->second.tokens.empty();  // committed already: free the shards; else abort the outstanding tokens
    op.tokens = std::move(it->second.tokens);
    op.copies = std::move(it->second.copies);
    sh.hook_queue.push_back(std::move(op));
  }
  sh.objects.erase(it);
  if (free_ranges) {
    allocator_->free_object(key);
    for (const auto& l : extra) allocator_->free_object(l);  // ledgers created by repair / demotion
  }
  if (was_complete) unpersist_object(sh, key);
  return ErrorCode::OK;
}

Result<bool> KeystoneService::object_exists(const ObjectKey& key) {
  if (key.empty()) return ErrorCode::INVALID_KEY;
  Shard& sh = shard_for(key);
  std::unique_lock<SpinMutex> lk(sh.mu);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end() || it->second.is_expired() || it->second.state != ObjectState::COMPLETE) return false;
  it->second.touch();
  return true;
}

Result<std::vector<CopyPlacement>> KeystoneService::get_workers(const ObjectKey& key) {
  if (key.empty()) return ErrorCode::INVALID_KEY;
  Shard& sh = shard_for(key);
  std::unique_lock<SpinMutex> lk(sh.mu);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end() || it->second.is_expired()) return ErrorCode::OBJECT_NOT_FOUND;
  if (it->second.state != ObjectState::COMPLETE) return ErrorCode::OBJECT_NOT_READY;
  if (it->second.copies.empty()) return ErrorCode::NO_COMPLETE_WORKER;
  it->second.touch();
  hot_.get_workers_total->fetch_add(1, std::memory_order_relaxed);
  if (config_.promote_after_reads > 0 && top_tier_rank_.load(std::memory_order_relaxed) >= 0) {
    int best = 1 << 20;
    for (const auto& c : it->second.copies)
      for (const auto& s : c.shards) best = std::min(best, tier_rank(s.storage_class));
    if (best > top_tier_rank_.load(std::memory_order_relaxed) &&
        ++it->second.reads_below_top == static_cast<uint32_t>(config_.promote_after_reads)) {
      std::lock_guard<std::mutex> pl(promo_mu_);
      promo_queue_.push_back(key);
    }
  }
  return it->second.copies;
}

Result<std::vector<CopyPlacement>> KeystoneService::put_start(const ObjectKey& key, size_t data_size, const WorkerConfig& config,
                                                              const std::string& client_id, const std::string& client_node) {
  const TimePoint t0 = Clock::now();
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  Result<std::vector<CopyPlacement>> r = ErrorCode::INTERNAL_ERROR;
  {
    std::shared_lock<std::shared_mutex> pk(pools_mu_);  // lock order: pools -> shard
    r = put_start_locked(key, data_size, config, client_id, client_node);
  }
  if (r.ok() && reservations_enabled()) reserve_after_start(key, r);
  if (r.ok()) {
    bump_view();
    hot_.put_start_total->fetch_add(1, std::memory_order_relaxed);
    hot_.put_bytes_total->fetch_add(data_size, std::memory_order_relaxed);
  } else {
    hot_.put_start_failed_total->fetch_add(1, std::memory_order_relaxed);
  }
  hot_.put_start_latency->observe(us_since(t0));
  return r;
}

// Caller holds pools_mu_ (shared).  No view bump / metrics here: batch_put_start does those once per batch.
Result<std::vector<CopyPlacement>> KeystoneService::put_start_locked(const ObjectKey& key, size_t data_size, const WorkerConfig& config,
                                                                     const std::string& client_id, const std::string& client_node) {
  if (key.empty() || key.find('\x01') != std::string::npos) return ErrorCode::INVALID_KEY;
  if (config.replication_factor == 0 || config.max_workers_per_copy == 0) return ErrorCode::INVALID_PARAMETERS;
  if (config_.max_replicas > 0 && config.replication_factor > static_cast<size_t>(config_.max_replicas)) return ErrorCode::VALUE_OUT_OF_RANGE;
  Shard& sh = shard_for(key);
  ShardGuard lk(this, sh);
  auto it = sh.objects.find(key);
  if (it != sh.objects.end()) {
    if (!it->second.is_expired()) return ErrorCode::OBJECT_ALREADY_EXISTS;
    erase_locked(sh, key, true);  // expired but not yet swept: reclaim inline
    metrics_.inc("expired_total");
  }
  WorkerConfig effective = config;
  if (effective.preferred_classes.empty() && !config_.tier_policy.empty()) {
    for (const auto& name : tier_classes_for_size(config_.tier_policy, data_size))
      if (auto sc = parse_storage_class(name)) effective.preferred_classes.push_back(*sc);
  }
  // admission: a tenant's put is checked against its grants and charged against its budget before anything is allocated
  const Tenant* ten = current_tenant();
  const uint64_t charge = static_cast<uint64_t>(data_size) * config.replication_factor;
  if (ten) {
    if (!ten->may_write(key)) {
      metrics_.inc("tenant_acl_denials_total");
      if (audit::enabled()) audit::event("acl_denied", {{"who", ten->name}, {"op", "write"}, {"key", key}});
      return ErrorCode::ACCESS_DENIED;
    }
    if (!tenant_admit(*ten, charge)) {
      metrics_.inc("tenant_quota_denials_total");
      if (audit::enabled()) audit::event("quota_denied", {{"who", ten->name}, {"key", key}, {"bytes", std::to_string(charge)}});
      return ErrorCode::QUOTA_EXCEEDED;
    }
  }
  auto copies = allocator_->allocate_data_copies(key, data_size, effective, pools_, client_node, draining_);  // (usually empty)
  if (!copies.ok()) {
    if (ten) tenant_release(ten->name, charge);
    return copies.error();
  }
  ObjectInfo info;
  info.key = key;
  info.size = data_size;
  info.created = info.last_accessed = Clock::now();
  info.config = config;
  for (auto& c : copies.value())
    for (auto& s : c.shards) s.checksum_algo = config.checksum;
  info.copies = copies.value();
  info.state = ObjectState::PENDING;
  info.owner_client = client_id;
  if (ten) info.tenant = ten->name;
  sh.objects.emplace(key, std::move(info));
  return copies;
}

ErrorCode KeystoneService::put_complete(const ObjectKey& key) { return put_complete(key, {}); }

ErrorCode KeystoneService::put_complete(const ObjectKey& key, const ShardChecksums& checksums) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  if (fault::fire("fail_put_complete")) return ErrorCode::INTERNAL_ERROR;
  Shard& sh = shard_for(key);
  if (reservations_enabled()) {
    // the tokens must still be valid at the workers: commit them BEFORE the object becomes readable.  A token that ran
    // out (OPERATION_TIMEOUT) means the worker has taken the range back -- the put failed, whatever the writer believes.
    std::vector<CopyPlacement> copies;
    ShardTokens tokens;
    TimePoint created;
    {
      std::lock_guard<SpinMutex> l0(sh.mu);
      auto i0 = sh.objects.find(key);
      if (i0 != sh.objects.end() && !i0->second.tokens.empty()) copies = i0->second.copies, tokens = i0->second.tokens, created = i0->second.created;
    }
    if (!tokens.empty()) {
      ReservationHooks hooks;
      {
        std::lock_guard<std::mutex> ml(mover_mu_);
        hooks = res_hooks_;
      }
      const ErrorCode cec = hooks.commit(copies, tokens);
      ShardGuard g(this, sh);
      auto i1 = sh.objects.find(key);
      const bool same = i1 != sh.objects.end() && i1->second.created == created;
      if (cec != ErrorCode::OK) {
        metrics_.inc("reservation_commit_failed_total");
        if (same) erase_locked(sh, key, true);  // aborts whatever is still reserved, frees the ledger
        return cec;
      }
      if (same) i1->second.tokens.clear();  // committed: from now on a removal frees shards instead of aborting tokens
    }
  }
  ShardGuard lk(this, sh);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end() || it->second.is_expired()) return ErrorCode::OBJECT_NOT_FOUND;
  ObjectInfo& info = it->second;
  if (!checksums.empty()) {
    if (checksums.size() != info.copies.size()) return ErrorCode::INVALID_PARAMETERS;
    for (size_t c = 0; c < checksums.size(); ++c) {
      if (checksums[c].size() != info.copies[c].shards.size()) return ErrorCode::INVALID_PARAMETERS;
      for (size_t s = 0; s < checksums[c].size(); ++s) info.copies[c].shards[s].checksum = checksums[c][s];
    }
  }
  if (!wal_enabled()) {  // volatile metadata: one critical section
    if (info.state != ObjectState::COMPLETE) {
      info.state = ObjectState::COMPLETE;
      hot_.put_complete_total->fetch_add(1, std::memory_order_relaxed);
    }
    info.touch();
    lk.finish();
    bump_view();
    return ErrorCode::OK;
  }
  // Durable before visible: the COMPLETE record is logged first (outside the spin lock); only when that succeeded does
  // the object become readable, and only then is the client told OK.  A failed log write leaves it PENDING.
  ErrorCode wal_ec = ErrorCode::OK;
  const TimePoint created = info.created;
  {
    const ObjectState prev = info.state;
    info.state = ObjectState::COMPLETE;
    persist_object(sh, info, &wal_ec);
    info.state = prev;
    info.committing = true;
  }
  lk.finish();  // writes the record, or waits for the thread that drained it
  lk.relock();
  it = sh.objects.find(key);
  const bool same = it != sh.objects.end() && it->second.created == created;
  if (same) it->second.committing = false;
  if (wal_ec != ErrorCode::OK) {
    lk.finish();
    metrics_.inc("wal_write_failed_total");
    return wal_ec == ErrorCode::NOT_LEADER ? wal_ec : ErrorCode::ETCD_ERROR;
  }
  if (!same) return ErrorCode::OBJECT_NOT_FOUND;  // cancelled / removed meanwhile (its tombstone is queued after our record)
  if (it->second.state != ObjectState::COMPLETE) {
    it->second.state = ObjectState::COMPLETE;
    hot_.put_complete_total->fetch_add(1, std::memory_order_relaxed);
  }
  it->second.touch();
  lk.finish();
  bump_view();
  return ErrorCode::OK;
}

ErrorCode KeystoneService::put_cancel(const ObjectKey& key) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  Shard& sh = shard_for(key);
  ShardGuard lk(this, sh);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end()) return ErrorCode::OBJECT_NOT_FOUND;
  if (it->second.state == ObjectState::COMPLETE) return ErrorCode::INVALID_STATE;
  erase_locked(sh, key, true);
  metrics_.inc("put_cancel_total");
  return ErrorCode::OK;
}

ErrorCode KeystoneService::remove_object(const ObjectKey& key) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  Shard& sh = shard_for(key);
  ShardGuard lk(this, sh);
  ErrorCode ec = erase_locked(sh, key, true);
  lk.finish();
  if (ec == ErrorCode::OK) {
    hot_.remove_total->fetch_add(1, std::memory_order_relaxed);
    bump_view();
  }
  return ec;
}

Result<size_t> KeystoneService::remove_all_objects() {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  size_t n = 0;
  for (auto& sh : shards_) {
    ShardGuard lk(this, sh);
    std::vector<ObjectKey> keys;
    keys.reserve(sh.objects.size());
    for (const auto& [k, v] : sh.objects) keys.push_back(k);
    for (const auto& k : keys) {
      erase_locked(sh, k, true);  // frees allocator ranges (the reference leaks them)
      ++n;
    }
  }
  bump_view();
  return n;
}

Result<ObjectInfo> KeystoneService::get_object_info(const ObjectKey& key) const {
  const Shard& sh = shard_for(key);
  std::shared_lock<SpinMutex> lk(sh.mu);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end()) return ErrorCode::OBJECT_NOT_FOUND;
  return it->second;
}

std::vector<KeystoneService::ListedObject> KeystoneService::list_objects(const std::string& prefix, size_t limit,
                                                                         const std::string& start_after) const {
  if (limit == 0 || limit > 10000) limit = 10000;
  const TimePoint now = Clock::now();
  std::vector<ListedObject> out;
  for (const auto& sh : shards_) {
    std::shared_lock<SpinMutex> lk(sh.mu);
    for (const auto& [k, o] : sh.objects) {
      if (o.state != ObjectState::COMPLETE || o.is_expired(now)) continue;
      if (k.compare(0, prefix.size(), prefix) != 0 || (!start_after.empty() && k <= start_after)) continue;
      ListedObject lo;
      lo.key = k;
      lo.size = o.size;
      lo.copies = static_cast<uint32_t>(o.copies.size());
      if (!o.copies.empty() && !o.copies[0].shards.empty()) lo.tier = o.copies[0].shards[0].storage_class;
      out.push_back(std::move(lo));
    }
  }
  // the table is hash-sharded: order globally, then cut the page (partial_sort keeps big listings cheap)
  const size_t n = std::min(limit, out.size());
  std::partial_sort(out.begin(), out.begin() + static_cast<std::ptrdiff_t>(n), out.end(), [](const ListedObject& a, const ListedObject& b) { return a.key < b.key; });
  out.resize(n);
  return out;
}

// ================================================================ batch API
std::vector<Result<bool>> KeystoneService::batch_object_exists(const std::vector<ObjectKey>& keys) {
  std::vector<Result<bool>> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(object_exists(k));
  return out;
}

std::vector<Result<std::vector<CopyPlacement>>> KeystoneService::batch_get_workers(const std::vector<ObjectKey>& keys) {
  std::vector<Result<std::vector<CopyPlacement>>> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(get_workers(k));
  return out;
}

bool KeystoneService::put_start_run(const std::vector<PutStartItem>& items, size_t first, size_t last, const std::string& client_id,
                                    const std::string& client_node, std::vector<Result<std::vector<CopyPlacement>>>& out) {
  const WorkerConfig& config = items[first].config;
  const size_t data_size = items[first].size;
  if (config.replication_factor != 1 || config.max_workers_per_copy == 0) return false;
  if (!draining_.empty()) return false;  // a worker is being drained: the per-object path knows which pools to leave alone
  if (current_tenant()) return false;    // a tenant's puts are admitted one by one against its grants and budget
  if (config_.max_replicas > 0 && config.replication_factor > static_cast<size_t>(config_.max_replicas)) return false;
  WorkerConfig tiered;
  const WorkerConfig* effective = &config;
  if (config.preferred_classes.empty() && !config_.tier_policy.empty()) {
    tiered = config;
    for (const auto& name : tier_classes_for_size(config_.tier_policy, data_size))
      if (auto sc = parse_storage_class(name)) tiered.preferred_classes.push_back(*sc);
    effective = &tiered;
  }
  std::vector<const ObjectKey*> keys;
  std::vector<size_t> index;  // keys[k] is items[index[k]]
  keys.reserve(last - first), index.reserve(last - first);
  for (size_t i = first; i < last; ++i) {
    const ObjectKey& key = items[i].key;
    if (key.empty() || key.find('\x01') != std::string::npos) {
      out[i] = ErrorCode::INVALID_KEY;
      continue;
    }
    keys.push_back(&key), index.push_back(i);
  }
  std::vector<alloc::IAllocator::RunSlot> slots;
  if (!keys.empty() && !allocator_->allocate_run(keys, data_size, *effective, pools_, client_node, slots)) return false;
  const TimePoint now = Clock::now();
  uint64_t placed_as_run = 0;
  for (size_t k = 0; k < keys.size(); ++k) {
    const size_t i = index[k];
    const ObjectKey& key = *keys[k];
    if (slots[k].status == ErrorCode::OBJECT_ALREADY_EXISTS) {
      // live object, expired-but-unswept one, or a racing put of the same key: the per-object path sorts out which
      out[i] = put_start_locked(key, data_size, config, client_id, client_node);
      continue;
    }
    if (slots[k].status != ErrorCode::OK) {
      out[i] = slots[k].status;
      continue;
    }
    slots[k].shard.checksum_algo = config.checksum;
    std::vector<CopyPlacement> copies(1);
    copies[0].copy_index = 0;
    copies[0].shards.push_back(std::move(slots[k].shard));
    bool clash = false;
    {
      Shard& sh = shard_for(key);
      ShardGuard lk(this, sh);
      if (sh.objects.count(key)) {
        clash = true;  // an object without a ledger entry (its pools were forgotten): give the extent back, then as above
      } else {
        ObjectInfo info;
        info.key = key;
        info.size = data_size;
        info.created = info.last_accessed = now;
        info.config = config;
        info.copies = copies;
        info.state = ObjectState::PENDING;
        info.owner_client = client_id;
        sh.objects.emplace(key, std::move(info));
      }
    }
    if (clash) {
      allocator_->free_object(key);
      out[i] = put_start_locked(key, data_size, config, client_id, client_node);
      continue;
    }
    out[i] = std::move(copies);
    ++placed_as_run;
  }
  if (placed_as_run) metrics_.inc("put_start_run_objects_total", placed_as_run);
  return true;
}

std::vector<Result<std::vector<CopyPlacement>>> KeystoneService::batch_put_start(const std::vector<PutStartItem>& items,
                                                                                 const std::string& client_id,
                                                                                 const std::string& client_node) {
  std::vector<Result<std::vector<CopyPlacement>>> out;
  out.reserve(items.size());
  if (!is_leader()) {
    out.assign(items.size(), ErrorCode::NOT_LEADER);
    return out;
  }
  const TimePoint t0 = Clock::now();
  uint64_t ok = 0, bytes = 0;
  {
    std::shared_lock<std::shared_mutex> pk(pools_mu_);  // once per batch
    // Runs of items with one size and one policy (a batch of activations, KV blocks, feature rows) are placed together
    constexpr size_t kMinRun = 8;
    size_t i = 0;
    while (i < items.size()) {
      size_t j = i + 1;
      while (j < items.size() && items[j].size == items[i].size && items[j].config == items[i].config) ++j;
      if (j - i >= kMinRun) {
        out.resize(j, ErrorCode::INTERNAL_ERROR);
        if (put_start_run(items, i, j, client_id, client_node, out)) {
          i = j;
          continue;
        }
        out.resize(i, ErrorCode::INTERNAL_ERROR);
      }
      for (; i < j; ++i) out.push_back(put_start_locked(items[i].key, items[i].size, items[i].config, client_id, client_node));
    }
  }
  if (reservations_enabled())
    for (size_t i = 0; i < items.size(); ++i)
      if (out[i].ok()) reserve_after_start(items[i].key, out[i]);
  for (size_t i = 0; i < items.size(); ++i)
    if (out[i].ok()) ++ok, bytes += items[i].size;
  if (ok) {
    bump_view();
    hot_.put_start_total->fetch_add(ok, std::memory_order_relaxed);
    hot_.put_bytes_total->fetch_add(bytes, std::memory_order_relaxed);
  }
  if (ok != items.size()) hot_.put_start_failed_total->fetch_add(items.size() - ok, std::memory_order_relaxed);
  if (!items.empty()) hot_.put_start_latency->observe(us_since(t0) / static_cast<double>(items.size()));
  return out;
}

std::vector<ErrorCode> KeystoneService::batch_put_complete(const std::vector<ObjectKey>& keys) {
  std::vector<ErrorCode> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(put_complete(k));
  return out;
}

std::vector<ErrorCode> KeystoneService::batch_put_complete(const std::vector<ObjectKey>& keys, const std::vector<ShardChecksums>& checksums) {
  std::vector<ErrorCode> out;
  out.reserve(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) out.push_back(put_complete(keys[i], i < checksums.size() ? checksums[i] : ShardChecksums{}));
  return out;
}

std::vector<ErrorCode> KeystoneService::batch_put_cancel(const std::vector<ObjectKey>& keys) {
  std::vector<ErrorCode> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(put_cancel(k));
  return out;
}

std::vector<ErrorCode> KeystoneService::batch_remove_object(const std::vector<ObjectKey>& keys) {
  std::vector<ErrorCode> out;
  out.reserve(keys.size());
  for (const auto& k : keys) out.push_back(remove_object(k));
  return out;
}

// ================================================================ cluster
Result<ClusterStats> KeystoneService::get_cluster_stats() const {
  ClusterStats st;
  {
    std::shared_lock<std::shared_mutex> lk(workers_mu_);
    st.total_workers = workers_.size();
  }
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    st.total_memory_pools = pools_.size();
    for (const auto& [id, p] : pools_) {
      st.total_capacity += p.size;
      st.used_capacity += allocator_->allocator().pool_used_bytes(id);
    }
  }
  for (const auto& sh : shards_) {
    std::shared_lock<SpinMutex> lk(sh.mu);
    for (const auto& [k, o] : sh.objects) (o.state == ObjectState::COMPLETE ? st.total_objects : st.pending_objects) += 1;
  }
  {
    std::lock_guard<std::mutex> lk(clients_mu_);
    st.active_clients = clients_.size();
  }
  st.avg_utilization = st.total_capacity ? static_cast<double>(st.used_capacity) / static_cast<double>(st.total_capacity) : 0.0;
  return st;
}

ErrorCode KeystoneService::get_workers_info(std::vector<WorkerInfo>& out) const {
  out.clear();
  std::shared_lock<std::shared_mutex> lk(workers_mu_);
  for (const auto& [id, w] : workers_) out.push_back(w);
  return ErrorCode::OK;
}

ErrorCode KeystoneService::get_memory_pools(std::vector<MemoryPool>& out) const {
  out.clear();
  std::shared_lock<std::shared_mutex> lk(pools_mu_);
  for (const auto& [id, p] : pools_) {
    out.push_back(p);
    out.back().used = allocator_->allocator().pool_used_bytes(id);
  }
  return ErrorCode::OK;
}

void KeystoneService::refresh_top_tier_locked() {
  int best = -1;
  for (const auto& [id, p] : pools_) {
    const int r = tier_rank(p.storage_class);
    if (best < 0 || r < best) best = r;
  }
  top_tier_rank_.store(best, std::memory_order_relaxed);
}

uint64_t KeystoneService::tier_capacity(StorageClass sc) const {
  uint64_t cap = 0;
  std::shared_lock<std::shared_mutex> lk(pools_mu_);
  for (const auto& [id, p] : pools_)
    if (p.storage_class == sc) cap += p.size;
  return cap;
}

double KeystoneService::tier_utilization(StorageClass sc) const {
  uint64_t cap = 0, used = 0;
  std::shared_lock<std::shared_mutex> lk(pools_mu_);
  for (const auto& [id, p] : pools_)
    if (p.storage_class == sc) {
      cap += p.size;
      used += allocator_->allocator().pool_used_bytes(id);
    }
  return cap ? static_cast<double>(used) / static_cast<double>(cap) : 0.0;
}

// ================================================================ client sessions
Result<std::string> KeystoneService::client_register(const std::string& node_id) {
  const std::string id = "client-" + uuid_to_string(generate_uuid()).substr(0, 16);
  std::lock_guard<std::mutex> lk(clients_mu_);
  clients_[id] = ClientSession{node_id, Clock::now()};
  return id;
}

Result<ViewVersionId> KeystoneService::client_ping(const std::string& client_id) {
  std::lock_guard<std::mutex> lk(clients_mu_);
  auto it = clients_.find(client_id);
  if (it == clients_.end()) return ErrorCode::SESSION_EXPIRED;
  it->second.last_ping = Clock::now();
  return get_view_version();
}

// ================================================================ GC / eviction / repair
void KeystoneService::set_copy_mover(CopyMover m) {
  std::lock_guard<std::mutex> lk(mover_mu_);
  mover_ = std::move(m);
}

void KeystoneService::set_copy_verifier(CopyVerifier v) {
  std::lock_guard<std::mutex> lk(mover_mu_);
  verifier_ = std::move(v);
}

void KeystoneService::set_reservation_hooks(ReservationHooks h) {
  std::lock_guard<std::mutex> lk(mover_mu_);
  reservations_on_.store(static_cast<bool>(h));
  res_hooks_ = std::move(h);
}

void KeystoneService::run_hook_ops(const std::vector<HookOp>& ops) {
  ReservationHooks hooks;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    hooks = res_hooks_;
  }
  if (!hooks) return;
  for (const auto& op : ops) {
    const ErrorCode ec = op.release ? hooks.release(op.copies) : hooks.abort(op.copies, op.tokens);
    if (ec != ErrorCode::OK) BB_VLOG(1) << "reservation " << (op.release ? "release" : "abort") << " at a worker failed: " << to_string(ec);
  }
}

// After put_start_locked created the PENDING object: reserve its shards at the workers (outside every lock).  On failure
// the object is withdrawn and `placed` carries the error.
ErrorCode KeystoneService::reserve_after_start(const ObjectKey& key, Result<std::vector<CopyPlacement>>& placed) {
  ReservationHooks hooks;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    hooks = res_hooks_;
  }
  ShardTokens tokens;
  const ErrorCode ec = hooks.reserve(key, placed.value(), static_cast<uint64_t>(std::max<int64_t>(1, config_.reservation_ttl_ms)), &tokens);
  Shard& sh = shard_for(key);
  ShardGuard lk(this, sh);
  auto it = sh.objects.find(key);
  if (ec != ErrorCode::OK) {
    metrics_.inc("reservation_failed_total");
    if (it != sh.objects.end() && it->second.state == ObjectState::PENDING) erase_locked(sh, key, true);
    placed = ec == ErrorCode::ALLOCATION_FAILED ? ErrorCode::INSUFFICIENT_SPACE : ec;
    return ec;
  }
  if (it != sh.objects.end()) {
    it->second.tokens = std::move(tokens);
    it->second.reserved = true;
  }
  metrics_.inc("reservations_total");
  return ErrorCode::OK;
}

// A worker reclaimed an expired reservation and says so: /.../reservations_expired/<worker>/<token> = <object key>.
void KeystoneService::on_reservation_expired(const std::string& key, const std::string& owner, bool is_delete) {
  if (is_delete || !is_leader()) return;
  const size_t slash = key.rfind('/');
  if (slash == std::string::npos) return;
  const std::string token = key.substr(slash + 1);
  bool dropped = false;
  {
    Shard& sh = shard_for(owner);
    ShardGuard lk(this, sh);
    auto it = sh.objects.find(owner);
    if (it != sh.objects.end() && it->second.state == ObjectState::PENDING) {
      bool mine = false;
      for (const auto& c : it->second.tokens)
        for (const auto& t : c) mine |= t == token;
      if (mine) {  // not a newer incarnation of the same key
        erase_locked(sh, owner, true);  // ledger freed; the object's OTHER tokens are aborted at their workers
        dropped = true;
      }
    }
  }
  if (dropped) {
    BB_LOG(INFO) << "keystone: writer of " << owner << " vanished (reservation " << token << " expired at its worker): put withdrawn";
    metrics_.inc("reservation_expired_total");
    bump_view();
  }
  if (coord_) coord_->del(key);
}

size_t KeystoneService::run_gc_once() {
  size_t n = 0;
  const TimePoint now = Clock::now();
  for (auto& sh : shards_) {
    ShardGuard lk(this, sh);
    std::vector<ObjectKey> dead;
    for (const auto& [k, o] : sh.objects) {
      if (o.is_expired(now)) dead.push_back(k);
      else if (o.state == ObjectState::PENDING && o.config.ttl_ms == 0 && now - o.created > kPendingGrace) dead.push_back(k);
    }
    for (const auto& k : dead) {
      erase_locked(sh, k, true);
      ++n;
    }
  }
  if (n) {
    metrics_.inc("expired_total", n);
    bump_view();
  }
  return n;
}

// Moves every copy of `key` to the first class of `targets` that has room (same replication factor), through the
// installed CopyMover (worker-to-worker D_COPY; fused kernel when the GPU tier is involved), then swaps the
// placements atomically and frees the old extents.  Used by watermark demotion and by explicit migrate_object().
ErrorCode KeystoneService::migrate_with(const CopyMover& mover, const ObjectKey& key, const std::vector<StorageClass>& targets,
                                        const PlacementFilter& accept) {
  auto info = get_object_info(key);
  if (!info.ok()) return info.error();
  if (info.value().copies.empty() || info.value().state != ObjectState::COMPLETE) return ErrorCode::OBJECT_NOT_READY;
  WorkerConfig cfg = info.value().config;
  cfg.replication_factor = info.value().copies.size();
  cfg.symmetric_replicas = false;
  std::string ledger;
  Result<std::vector<CopyPlacement>> fresh = ErrorCode::INSUFFICIENT_SPACE;
  // one rung at a time: the first target tier that has room (GPU -> DRAM before GPU -> NVMe)
  for (StorageClass target : targets) {
    cfg.preferred_classes = {target};
    for (int slot = 0; slot < 64 && !fresh.ok(); ++slot) {
      ledger = key + "\x01" + std::to_string(slot);
      std::shared_lock<std::shared_mutex> pk(pools_mu_);
      alloc::IAllocator::PoolMap eligible;
      for (const auto& [pid, p] : pools_)
        if (p.storage_class == target && std::find(draining_.begin(), draining_.end(), pid) == draining_.end()) eligible.emplace(pid, p);
      if (eligible.empty()) break;
      fresh = allocator_->allocate_data_copies(ledger, info.value().size, cfg, eligible);
      if (!fresh.ok() && fresh.error() != ErrorCode::OBJECT_ALREADY_EXISTS) break;
    }
    if (fresh.ok()) break;
  }
  if (!fresh.ok()) return fresh.error();
  if (accept && !accept(info.value().copies, fresh.value())) {
    allocator_->free_object(ledger);
    return ErrorCode::INVALID_STATE;  // the caller's filter turned the new placement down; nothing moved
  }
  ErrorCode ec = ErrorCode::OK;
  for (size_t c = 0; c < fresh.value().size() && ec == ErrorCode::OK; ++c)
    ec = mover(key, info.value().copies[std::min(c, info.value().copies.size() - 1)], fresh.value()[c], info.value().config.checksum);
  if (ec != ErrorCode::OK) {
    allocator_->free_object(ledger);
    return ec;
  }
  Shard& sh = shard_for(key);
  ShardGuard lk(this, sh);
  auto it = sh.objects.find(key);
  if (it == sh.objects.end() || it->second.created != info.value().created) {  // removed / replaced while we copied
    lk.finish();
    allocator_->free_object(ledger);
    return ErrorCode::OBJECT_NOT_FOUND;
  }
  it->second.copies = fresh.value();
  const std::vector<std::string> old = std::move(it->second.extra_ledgers);
  it->second.extra_ledgers = {ledger};
  persist_object(sh, it->second);
  lk.finish();
  allocator_->free_object(key);  // old extents
  for (const auto& l : old) allocator_->free_object(l);
  bump_view();
  return ErrorCode::OK;
}

namespace {
// Highest end offset (pool-relative) of an object's shards inside `pool`; 0 when it has none there.
uint64_t top_in_pool(const std::vector<CopyPlacement>& copies, const MemoryPool& pool) {
  uint64_t top = 0;
  const uint64_t base = pool.ucx_remote_addr ? pool.ucx_remote_addr : pool.base_addr;
  for (const auto& c : copies)
    for (const auto& s : c.shards) {
      if (s.pool_id != pool.id) continue;
      uint64_t off = 0;
      if (auto* g = std::get_if<GpuSlabLocation>(&s.location)) off = g->offset;
      else if (auto* f = std::get_if<FileLocation>(&s.location)) off = f->file_offset;
      else if (auto* x = std::get_if<CxlMemoryLocation>(&s.location)) off = x->offset;
      else if (auto* m = std::get_if<MemoryLocation>(&s.location)) off = m->remote_addr >= base ? m->remote_addr - base : 0;
      top = std::max(top, off + s.length);
    }
  return top;
}
}  // namespace

Result<size_t> KeystoneService::compact_pool(const MemoryPoolId& pool_id, size_t max_moves) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  if (!mover) return ErrorCode::NOT_IMPLEMENTED;
  MemoryPool pool;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    auto it = pools_.find(pool_id);
    if (it == pools_.end()) return ErrorCode::MEMORY_POOL_NOT_FOUND;
    pool = it->second;
  }
  size_t moved = 0;
  std::unordered_set<ObjectKey> tried;  // an object whose move was turned down is not retried in this run
  while (moved < max_moves) {
    // the not-yet-tried object that reaches highest into the pool
    ObjectKey victim;
    uint64_t victim_top = 0;
    for (const auto& sh : shards_) {
      std::shared_lock<SpinMutex> lk(sh.mu);
      for (const auto& [k, o] : sh.objects) {
        if (o.state != ObjectState::COMPLETE || tried.count(k)) continue;
        const uint64_t top = top_in_pool(o.copies, pool);
        if (top > victim_top) victim_top = top, victim = k;
      }
    }
    if (victim.empty()) break;
    tried.insert(victim);
    // new home: same tier, and strictly lower inside this pool (or out of it) -- otherwise the move is pointless
    const ErrorCode ec = migrate_with(mover, victim, {pool.storage_class},
                                      [&](const std::vector<CopyPlacement>& old_copies, const std::vector<CopyPlacement>& fresh) {
                                        return top_in_pool(fresh, pool) < top_in_pool(old_copies, pool);
                                      });
    if (ec == ErrorCode::OK) {
      ++moved;
      tried.erase(victim);  // it may move again, further down, once other holes open
      metrics_.inc("compaction_moves_total");
    }
  }
  if (moved) bump_view();
  return moved;
}

size_t KeystoneService::run_compaction_once() {
  if (config_.compaction_fragmentation_threshold <= 0.0 || !is_leader()) return 0;
  std::vector<MemoryPoolId> fragmented;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    for (const auto& [id, p] : pools_)
      if (allocator_->allocator().pool_fragmentation(id) > config_.compaction_fragmentation_threshold) fragmented.push_back(id);
  }
  size_t moved = 0;
  for (const auto& id : fragmented) {
    auto r = compact_pool(id, 8);
    if (r.ok()) moved += r.value();
  }
  return moved;
}

ErrorCode KeystoneService::migrate_object(const ObjectKey& key, StorageClass target) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  if (!mover) return ErrorCode::NOT_IMPLEMENTED;
  {
    auto info = get_object_info(key);
    if (!info.ok()) return info.error();
    bool already = !info.value().copies.empty();
    for (const auto& c : info.value().copies)
      for (const auto& s : c.shards) already &= s.storage_class == target;
    if (already) return ErrorCode::OK;
  }
  ErrorCode ec = migrate_with(mover, key, {target});
  if (ec == ErrorCode::OK) metrics_.inc("migrations_total");
  return ec;
}

size_t KeystoneService::run_promotion_once() {
  std::vector<ObjectKey> keys;
  {
    std::lock_guard<std::mutex> lk(promo_mu_);
    keys.swap(promo_queue_);
  }
  if (keys.empty() || !is_leader()) return 0;
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  if (!mover) return 0;
  // tiers present in the cluster, fastest first
  std::map<int, std::vector<StorageClass>> tiers;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    for (const auto& [id, p] : pools_) {
      auto& v = tiers[tier_rank(p.storage_class)];
      if (std::find(v.begin(), v.end(), p.storage_class) == v.end()) v.push_back(p.storage_class);
    }
  }
  size_t promoted = 0;
  for (const auto& key : keys) {
    auto info = get_object_info(key);
    if (!info.ok() || info.value().copies.empty()) continue;
    int cur = 1 << 20;
    for (const auto& c : info.value().copies)
      for (const auto& s : c.shards) cur = std::min(cur, tier_rank(s.storage_class));
    std::vector<StorageClass> targets;  // faster tiers that would stay under the watermark with this object added
    for (const auto& [rank, classes] : tiers) {
      if (rank >= cur) break;
      for (StorageClass sc : classes) {
        const double cap = static_cast<double>(tier_capacity(sc));
        if (cap > 0 && tier_utilization(sc) + static_cast<double>(info.value().size * info.value().copies.size()) / cap < config_.high_watermark)
          targets.push_back(sc);
      }
    }
    if (targets.empty()) continue;
    if (migrate_with(mover, key, targets) == ErrorCode::OK) {
      ++promoted;
      Shard& sh = shard_for(key);
      std::unique_lock<SpinMutex> lk(sh.mu);
      auto it = sh.objects.find(key);
      if (it != sh.objects.end()) it->second.reads_below_top = 0;
    }
  }
  if (promoted) metrics_.inc("promotions_total", promoted);
  return promoted;
}

size_t KeystoneService::run_eviction_once() {
  // utilisation per tier from the allocator's live accounting
  std::map<int, std::vector<StorageClass>> tiers;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    std::set<StorageClass> seen;
    for (const auto& [id, p] : pools_) seen.insert(p.storage_class);
    for (auto sc : seen) tiers[tier_rank(sc)].push_back(sc);
  }
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  size_t total = 0;
  for (const auto& [rank, classes] : tiers) {
    for (StorageClass sc : classes) {
      if (tier_utilization(sc) < config_.high_watermark) continue;
      // LRU candidates that have data on this tier and are not soft-pinned
      std::vector<std::pair<TimePoint, ObjectKey>> cands;
      for (auto& sh : shards_) {
        std::shared_lock<SpinMutex> lk(sh.mu);
        for (const auto& [k, o] : sh.objects) {
          if (o.state != ObjectState::COMPLETE || o.config.enable_soft_pin) continue;
          bool here = false;
          for (const auto& c : o.copies)
            for (const auto& s : c.shards) here |= s.storage_class == sc;
          if (here) cands.emplace_back(o.last_accessed, k);
        }
      }
      if (cands.empty()) continue;
      std::sort(cands.begin(), cands.end());
      const size_t n = std::max<size_t>(1, static_cast<size_t>(static_cast<double>(cands.size()) * config_.eviction_ratio));
      // lower tiers that exist in the cluster
      std::vector<StorageClass> lower;
      for (const auto& [r2, cl2] : tiers)
        if (r2 > rank) lower.insert(lower.end(), cl2.begin(), cl2.end());
      for (size_t i = 0; i < n && i < cands.size(); ++i) {
        const ObjectKey& key = cands[i].second;
        const bool demoted = mover && !lower.empty() && migrate_with(mover, key, lower) == ErrorCode::OK;
        if (demoted) metrics_.inc("demotions_total");
        if (!demoted) {
          Shard& sh = shard_for(key);
          ShardGuard lk(this, sh);
          if (erase_locked(sh, key, true) == ErrorCode::OK) metrics_.inc("evictions_total");
        }
        ++total;
      }
    }
  }
  if (total) bump_view();
  return total;
}

size_t KeystoneService::run_repair_once() {
  CopyMover mover;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
  }
  if (!mover) return 0;
  std::vector<ObjectInfo> degraded;
  for (auto& sh : shards_) {
    std::shared_lock<SpinMutex> lk(sh.mu);
    for (const auto& [k, o] : sh.objects)
      if (o.state == ObjectState::COMPLETE && !o.copies.empty() && o.copies.size() < o.config.replication_factor) degraded.push_back(o);
  }
  size_t repaired = 0;
  for (const auto& o : degraded) {
    std::string ledger;
    auto placed = place_extra_copy(o, o.copies.size(), ledger);
    if (!placed.ok()) continue;
    CopyPlacement dst = placed.value();
    // any surviving copy can be the source: one that fails its digest on the way (bit rot) must not block the repair
    ErrorCode mec = ErrorCode::OBJECT_NOT_FOUND;
    for (size_t c = 0; c < o.copies.size() && mec != ErrorCode::OK; ++c) mec = mover(o.key, o.copies[c], dst, o.config.checksum);
    if (mec != ErrorCode::OK) {
      allocator_->free_object(ledger);
      continue;
    }
    Shard& sh = shard_for(o.key);
    ShardGuard lk(this, sh);
    auto it = sh.objects.find(o.key);
    if (it == sh.objects.end() || it->second.created != o.created) {
      lk.finish();
      allocator_->free_object(ledger);
      continue;
    }
    dst.copy_index = static_cast<uint32_t>(it->second.copies.size());
    it->second.copies.push_back(dst);
    it->second.extra_ledgers.push_back(ledger);
    persist_object(sh, it->second);
    ++repaired;
  }
  if (repaired) {
    metrics_.inc("repairs_total", repaired);
    bump_view();
  }
  return repaired;
}

Result<ScrubReport> KeystoneService::scrub(const std::string& prefix, size_t max_objects) {
  if (!is_leader()) return ErrorCode::NOT_LEADER;
  size_t cursor = 0;
  return scrub_from(prefix, max_objects, cursor);
}

Result<ScrubReport> KeystoneService::scrub_from(const std::string& prefix, size_t max_objects, size_t& cursor) {
  CopyMover mover;
  CopyVerifier verify;
  {
    std::lock_guard<std::mutex> lk(mover_mu_);
    mover = mover_;
    verify = verifier_;
  }
  if (!verify) return ErrorCode::NOT_IMPLEMENTED;
  ScrubReport rep;
  // whole metadata shards at a time, starting where the last background round stopped
  for (size_t n = 0; n < shards_.size() && (max_objects == 0 || rep.objects < max_objects); ++n, cursor = (cursor + 1) % shards_.size()) {
    std::vector<ObjectInfo> batch;
    {
      Shard& sh = shards_[cursor % shards_.size()];
      std::shared_lock<SpinMutex> lk(sh.mu);
      for (const auto& [k, o] : sh.objects)
        if (o.state == ObjectState::COMPLETE && !o.copies.empty() && k.compare(0, prefix.size(), prefix) == 0) batch.push_back(o);
    }
    for (const auto& o : batch) {
      ++rep.objects;
      const ChecksumAlgo algo = o.config.checksum;
      std::vector<size_t> good, bad;
      for (size_t c = 0; c < o.copies.size(); ++c) {
        ++rep.copies;
        const ErrorCode ec = verify(o.key, o.copies[c], algo);
        if (ec == ErrorCode::OK) good.push_back(c);
        else if (ec == ErrorCode::CHECKSUM_MISMATCH) bad.push_back(c);
        else ++rep.unreachable;
      }
      if (bad.empty()) continue;
      {
        // hashed from a snapshot: an object that was removed, replaced or moved meanwhile had its extents reused under the
        // hash -- that is a lost race, not rot, and the new incarnation is looked at on the next pass
        Shard& sh = shard_for(o.key);
        std::shared_lock<SpinMutex> lk(sh.mu);
        auto it = sh.objects.find(o.key);
        if (it == sh.objects.end() || it->second.created != o.created || !(it->second.copies == o.copies)) continue;
      }
      rep.corrupt += bad.size();
      metrics_.inc("scrub_corrupt_copies_total", bad.size());
      if (good.empty()) {
        if (bad.size() == o.copies.size()) {
          ++rep.unrecoverable;
          metrics_.inc("scrub_unrecoverable_total");
          BB_LOG(ERROR) << "scrub: " << o.key << " has no copy left that matches its digest";
        }
        continue;  // (a copy that could not be asked may still be fine: try again next round)
      }
      if (!mover) continue;
      for (size_t c : bad)
        if (replace_copy(mover, o, c, o.copies[good[0]]) == ErrorCode::OK) {
          ++rep.healed;
          metrics_.inc("scrub_healed_total");
        }
    }
  }
  metrics_.inc("scrub_objects_total", rep.objects);
  if (rep.healed) bump_view();
  return rep;
}

// Swaps copy `idx` of `o` for a fresh one made from `source` (a copy of the same object that verified): new extents under
// their own ledger entry, bytes moved and digested by the mover, metadata switched under the shard lock if the object is
// still the same incarnation, and only then the bad extents handed back to the allocator.
// Extents for one more copy of `o` (repair) or for the successor of copy `skip` (scrub, drain): never a pool that holds
// another replica, and -- unless the cluster is too small -- no pool of a worker that does (a worker death must not take two
// copies).  `ledger` receives the allocator entry the extents are booked under.
Result<CopyPlacement> KeystoneService::place_extra_copy(const ObjectInfo& o, size_t skip, std::string& ledger) {
  WorkerConfig cfg = o.config;
  cfg.replication_factor = 1;
  cfg.symmetric_replicas = false;
  // a successor belongs on the tier the copy is on now (an object that was demoted stays demoted); the put-time
  // preferences follow as fall-backs
  if (skip < o.copies.size() && !o.copies[skip].shards.empty()) {
    const StorageClass here = o.copies[skip].shards[0].storage_class;
    cfg.preferred_classes.erase(std::remove(cfg.preferred_classes.begin(), cfg.preferred_classes.end(), here), cfg.preferred_classes.end());
    cfg.preferred_classes.insert(cfg.preferred_classes.begin(), here);
  }
  Result<std::vector<CopyPlacement>> fresh = ErrorCode::INSUFFICIENT_SPACE;
  for (int pass = 0; pass < 2 && !fresh.ok(); ++pass) {
    std::shared_lock<std::shared_mutex> pk(pools_mu_);
    std::vector<MemoryPoolId> exclude;
    std::unordered_set<WorkerId> busy;
    for (size_t c = 0; c < o.copies.size(); ++c)
      if (c != skip)
        for (const auto& s : o.copies[c].shards) {
          exclude.push_back(s.pool_id);
          busy.insert(s.worker_id);
        }
    alloc::IAllocator::PoolMap eligible;
    for (const auto& [pid, p] : pools_) {
      if (std::find(draining_.begin(), draining_.end(), pid) != draining_.end()) continue;
      if (pass == 0 && busy.count(p.worker_id)) continue;
      eligible.emplace(pid, p);
    }
    if (eligible.empty()) continue;
    for (int slot = 0; slot < 64; ++slot) {
      ledger = o.key + "\x01" + std::to_string(slot);
      fresh = allocator_->allocate_data_copies(ledger, o.size, cfg, eligible, "", exclude);
      if (fresh.ok() || fresh.error() != ErrorCode::OBJECT_ALREADY_EXISTS) break;
    }
  }
  if (!fresh.ok()) return fresh.error();
  return fresh.value()[0];
}

ErrorCode KeystoneService::replace_copy(const CopyMover& mover, const ObjectInfo& o, size_t idx, const CopyPlacement& source) {
  std::string ledger;
  auto placed = place_extra_copy(o, idx, ledger);
  if (!placed.ok()) return placed.error();
  CopyPlacement dst = placed.value();
  const ErrorCode mec = mover(o.key, source, dst, o.config.checksum);
  if (mec != ErrorCode::OK) {
    allocator_->free_object(ledger);
    return mec;
  }
  std::vector<std::string> ledgers;
  {
    Shard& sh = shard_for(o.key);
    ShardGuard lk(this, sh);
    auto it = sh.objects.find(o.key);
    // the copy list may have changed under us (migration, repair, a pool that left): only swap what we looked at
    if (it == sh.objects.end() || it->second.created != o.created || idx >= it->second.copies.size() || !(it->second.copies[idx] == o.copies[idx])) {
      lk.finish();
      allocator_->free_object(ledger);
      return ErrorCode::OBJECT_NOT_FOUND;
    }
    dst.copy_index = it->second.copies[idx].copy_index;
    it->second.copies[idx] = dst;
    it->second.extra_ledgers.push_back(ledger);
    ledgers = it->second.extra_ledgers;
    persist_object(sh, it->second);
  }
  // hand the bad extents back: they sit in the object's own ledger entry or in one a repair / migration made
  alloc::IAllocator::PoolMap pools;
  {
    std::shared_lock<std::shared_mutex> pk(pools_mu_);
    for (const auto& s : o.copies[idx].shards) {
      auto pit = pools_.find(s.pool_id);
      if (pit != pools_.end()) pools.emplace(pit->first, pit->second);
    }
  }
  size_t freed = allocator_->allocator().free_extents(o.key, o.copies[idx].shards, pools);
  for (const auto& l : ledgers)
    if (freed < o.copies[idx].shards.size() && l != ledger) freed += allocator_->allocator().free_extents(l, o.copies[idx].shards, pools);
  return ErrorCode::OK;
}

// ================================================================ tenants (admission control)
bool KeystoneService::tenant_admit(const Tenant& t, uint64_t bytes) {
  std::lock_guard<std::mutex> lk(tenant_mu_);
  TenantCount& c = tenant_usage_[t.name];
  if (t.quota_bytes && (bytes > t.quota_bytes || c.bytes > t.quota_bytes - bytes)) return false;
  if (t.max_objects && c.objects >= t.max_objects) return false;
  c.bytes += bytes;
  c.objects += 1;
  return true;
}

void KeystoneService::tenant_charge(const std::string& name, uint64_t bytes) {
  std::lock_guard<std::mutex> lk(tenant_mu_);
  TenantCount& c = tenant_usage_[name];
  c.bytes += bytes;
  c.objects += 1;
}

void KeystoneService::tenant_release(const std::string& name, uint64_t bytes) {
  std::lock_guard<std::mutex> lk(tenant_mu_);
  auto it = tenant_usage_.find(name);
  if (it == tenant_usage_.end()) return;
  it->second.bytes -= std::min(it->second.bytes, bytes);
  if (it->second.objects) --it->second.objects;
}

std::vector<TenantUsage> KeystoneService::tenant_usage() const {
  std::map<std::string, TenantUsage> by_name;
  for (const auto& n : tenant_names()) by_name[n].name = n;
  {
    std::lock_guard<std::mutex> lk(tenant_mu_);
    for (const auto& [n, c] : tenant_usage_) {
      if (!c.objects && !c.bytes && !by_name.count(n)) continue;
      TenantUsage& u = by_name[n];
      u.name = n;
      u.used_bytes = c.bytes;
      u.objects = c.objects;
    }
  }
  std::vector<TenantUsage> out;
  for (auto& [n, u] : by_name) {
    if (auto t = find_tenant(n)) u.quota_bytes = t->quota_bytes, u.max_objects = t->max_objects;
    out.push_back(std::move(u));
  }
  return out;
}

// ================================================================ metadata log
void KeystoneService::persist_object(Shard& sh, const ObjectInfo& info, ErrorCode* result) {
  if (!wal_enabled()) return;
  sh.wal_queue.push_back(WalOp{info.key, encode_object(info), false, result});
}

void KeystoneService::unpersist_object(Shard& sh, const ObjectKey& key) {
  if (!wal_enabled()) return;
  sh.wal_queue.push_back(WalOp{key, {}, true, nullptr});
}

// Writes the shard's queued records in queue order.  Whoever holds wal_mu writes everything queued so far, so a thread
// that finds the queue already drained only had to wait for the writer: when flush_wal returns, every record enqueued
// by the caller before it is on the log (and `result` slots are filled).
void KeystoneService::flush_wal(Shard& sh) {
  std::lock_guard<std::mutex> wl(sh.wal_mu);
  std::vector<WalOp> ops;
  {
    std::lock_guard<SpinMutex> lk(sh.mu);
    ops.swap(sh.wal_queue);
  }
  if (ops.empty()) return;
  uint64_t last_seq = 0;
  for (const auto& op : ops) {
    ErrorCode ec;
    if (local_wal_) {
      wire::Writer w;
      w.u8(op.del ? 2 : 1);
      w.str(op.key);
      if (!op.del) w.str(op.value);
      const uint64_t seq = local_wal_->append(w.data());
      ec = seq ? ErrorCode::OK : ErrorCode::IO_ERROR;
      if (seq) last_seq = seq;
    } else {
      ec = wal_write(op);
    }
    if (op.result) *op.result = ec;
  }
  if (local_wal_ && last_seq) {
    const ErrorCode sc = local_wal_->sync(last_seq);  // one fdatasync for the whole batch (group commit)
    if (sc != ErrorCode::OK)
      for (const auto& op : ops)
        if (op.result && *op.result == ErrorCode::OK) *op.result = sc;
    if (local_wal_->snapshot_due()) maybe_snapshot_local_wal();
  }
}

// Coordination-store sink: every write is a transaction guarded by the election key's create revision (the term).
ErrorCode KeystoneService::wal_write(const WalOp& op) {
  if (!coord_ || !coord_->is_connected()) return ErrorCode::ETCD_ERROR;
  const std::string k = cluster_prefix() + "objects/" + op.key;
  const ErrorCode ec = op.del ? coord_->fenced_del(election_name(), k) : coord_->fenced_put(election_name(), k, op.value);
  if (ec == ErrorCode::NOT_LEADER) step_down("a fenced metadata write was refused (another leader holds a newer term)");
  return ec;
}

void KeystoneService::recover_objects_from_wal() {
  if (!coord_ || !coord_->is_connected()) return;
  std::vector<std::string> keys, values;
  const std::string pfx = cluster_prefix() + "objects/";
  if (coord_->get_with_prefix(pfx, keys, values) != ErrorCode::OK) return;
  std::vector<std::pair<std::string, std::string>> recs;
  recs.reserve(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) recs.emplace_back(keys[i].substr(std::min(pfx.size(), keys[i].size())), std::move(values[i]));
  recover_objects(recs);
}

// Rebuilds table + allocator reservations from (key, encoded ObjectInfo) records.  The table must be empty of these keys.
void KeystoneService::recover_objects(std::vector<std::pair<std::string, std::string>>& records) {
  size_t n = 0, deferred = 0;
  std::shared_lock<std::shared_mutex> pk(pools_mu_);
  auto& ra = static_cast<alloc::RangeAllocator&>(allocator_->allocator());
  for (auto& [k, v] : records) {
    ObjectInfo o;
    if (!decode_object(v, o) || o.is_expired()) continue;
    o.state = ObjectState::COMPLETE;
    Shard& sh = shard_for(o.key);
    std::unique_lock<SpinMutex> lk(sh.mu);
    if (sh.objects.count(o.key)) continue;
    // re-reserve the extents so that new puts cannot overwrite recovered objects; shards whose pool has not
    // registered yet are adopted when it does (register_memory_pool)
    ra.adopt(o.key, o.copies, pools_);
    for (const auto& c : o.copies)
      for (const auto& s : c.shards)
        if (!pools_.count(s.pool_id)) {
          std::lock_guard<std::mutex> ul(unadopted_mu_);
          auto& v2 = unadopted_[s.pool_id];
          if (v2.empty() || v2.back() != o.key) v2.push_back(o.key), ++deferred;
        }
    if (!o.tenant.empty()) tenant_charge(o.tenant, tenant_charge_of(o));
    sh.objects.emplace(o.key, std::move(o));
    ++n;
  }
  if (n) BB_LOG(INFO) << "keystone: recovered " << n << " objects from the metadata log" << (deferred ? " (" + std::to_string(deferred) + " await their pools)" : "");
  metrics_.inc("objects_recovered_total", n);
}

// Local sink (`wal_path`): snapshot + log in a directory of this host.
ErrorCode KeystoneService::open_local_wal() {
  auto log = std::make_unique<DurableLog>();
  DurableLog::Options o;
  o.dir = config_.wal_path;
  o.name = "keystone-" + config_.cluster_id;
  o.fsync = config_.wal_fsync;
  o.snapshot_bytes = config_.wal_snapshot_mb > 0 ? static_cast<uint64_t>(config_.wal_snapshot_mb) << 20 : (64ull << 20);
  std::string snap;
  std::unordered_map<std::string, std::string> live;
  std::vector<std::string> order;
  auto put = [&](std::string k, std::string v) {
    auto [it, fresh] = live.try_emplace(k, std::move(v));
    if (fresh) order.push_back(std::move(k));
    else it->second = std::move(v);
  };
  // records first land in `pending`: the snapshot must be applied before them
  std::vector<std::string> pending;
  ErrorCode ec = log->open(o, &snap, [&](std::string_view r) { pending.emplace_back(r); });
  if (ec != ErrorCode::OK) {
    BB_LOG(ERROR) << "keystone: cannot open the metadata log under " << config_.wal_path;
    return ec;
  }
  if (!snap.empty()) {
    wire::Reader r(snap);
    const uint32_t n = r.u32();
    for (uint32_t i = 0; i < n && r.ok(); ++i) {
      std::string k = r.str(), v = r.str();
      if (r.ok()) put(std::move(k), std::move(v));
    }
  }
  for (const auto& rec : pending) {
    wire::Reader r(rec.data(), rec.size());
    const uint8_t op = r.u8();
    std::string k = r.str();
    if (op == 1) {
      std::string v = r.str();
      if (r.ok()) put(std::move(k), std::move(v));
    } else if (op == 2 && r.ok()) {
      live.erase(k);
    }
  }
  std::vector<std::pair<std::string, std::string>> recs;
  for (auto& k : order) {
    auto it = live.find(k);
    if (it != live.end()) recs.emplace_back(k, std::move(it->second));
  }
  recover_objects(recs);
  local_wal_ = std::move(log);
  return ErrorCode::OK;
}

void KeystoneService::maybe_snapshot_local_wal() {
  if (!local_wal_ || !local_wal_->is_open()) return;
  std::unique_lock<std::mutex> sl(local_snap_mu_, std::try_to_lock);
  if (!sl.owns_lock()) return;
  // Records are whole-object PUTs / tombstones, applied to memory before they are logged (put_complete's COMPLETE
  // record is covered by `committing`), so a state capture taken after rotate() contains every mutation of the old
  // generations, and replaying the new generation on top of it is idempotent: no global lock is needed.
  const uint64_t gen = local_wal_->rotate();
  wire::Writer w;
  std::vector<std::pair<std::string, std::string>> objs;
  for (auto& sh : shards_) {
    std::lock_guard<SpinMutex> lk(sh.mu);
    for (const auto& [k, o] : sh.objects) {
      if (o.state != ObjectState::COMPLETE && !o.committing) continue;
      ObjectInfo c = o;
      c.state = ObjectState::COMPLETE;
      objs.emplace_back(k, encode_object(c));
    }
  }
  w.u32(static_cast<uint32_t>(objs.size()));
  for (const auto& [k, v] : objs) {
    w.str(k);
    w.str(v);
  }
  if (local_wal_->install_snapshot(gen, w.data()) != ErrorCode::OK) BB_LOG(ERROR) << "keystone: metadata snapshot failed";
  else metrics_.inc("wal_snapshots_total");
}

// ================================================================ observability
std::string KeystoneService::metrics_text() const {
  auto st = get_cluster_stats();
  if (st.ok()) {
    metrics_.set_gauge("objects", static_cast<double>(st.value().total_objects));
    metrics_.set_gauge("pending_objects", static_cast<double>(st.value().pending_objects));
    metrics_.set_gauge("workers", static_cast<double>(st.value().total_workers));
    metrics_.set_gauge("memory_pools", static_cast<double>(st.value().total_memory_pools));
    metrics_.set_gauge("capacity_bytes", static_cast<double>(st.value().total_capacity));
    metrics_.set_gauge("used_bytes", static_cast<double>(st.value().used_capacity));
    metrics_.set_gauge("utilization", st.value().avg_utilization);
    metrics_.set_gauge("active_clients", static_cast<double>(st.value().active_clients));
  }
  metrics_.set_gauge("view_version", static_cast<double>(get_view_version()));
  metrics_.set_gauge("is_leader", is_leader() ? 1.0 : 0.0);
  const auto as = allocator_->get_allocator_stats();
  metrics_.set_gauge("allocator_fragmentation", as.fragmentation_ratio);
  metrics_.set_gauge("allocator_shards", static_cast<double>(as.total_shards));
  std::string out = metrics_.render("bb_");
  // per-tier gauges with labels
  std::map<StorageClass, std::pair<uint64_t, uint64_t>> tiers;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    for (const auto& [id, p] : pools_) {
      tiers[p.storage_class].first += p.size;
      tiers[p.storage_class].second += allocator_->allocator().pool_used_bytes(id);
    }
  }
  out += "# TYPE bb_tier_capacity_bytes gauge\n";
  for (const auto& [sc, v] : tiers) out += "bb_tier_capacity_bytes{tier=\"" + std::string(to_string(sc)) + "\"} " + std::to_string(v.first) + "\n";
  out += "# TYPE bb_tier_used_bytes gauge\n";
  for (const auto& [sc, v] : tiers) out += "bb_tier_used_bytes{tier=\"" + std::string(to_string(sc)) + "\"} " + std::to_string(v.second) + "\n";
  return out;
}

Json KeystoneService::stats_json() const {
  Json j = Json::object();
  auto st = get_cluster_stats();
  if (st.ok()) j["cluster"] = to_json(st.value());
  j["view_version"] = get_view_version();
  j["is_leader"] = is_leader();
  j["service_id"] = config_.service_id;
  j["cluster_id"] = config_.cluster_id;
  Json pools = Json::array();
  std::vector<MemoryPool> mp;
  get_memory_pools(mp);
  for (const auto& p : mp) pools.push_back(to_json(p));
  j["pools"] = pools;
  return j;
}

}  // namespace bb::keystone
