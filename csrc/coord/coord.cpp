#include "coord/coord.h"
#include "coord/etcd_coord.h"

#include <algorithm>
#include <chrono>
#include <sstream>

#include "common/log.h"
#include "rpc/wire.h"

namespace bb::coord {

namespace {
enum LogOp : uint32_t { OP_PUT = 1, OP_DEL = 2, OP_GRANT = 3, OP_REVOKE = 4 };
}

// ================================================================ MemCoord
MemCoord::MemCoord() {
  expiry_thread_ = std::thread([this] { expiry_loop(); });
  dispatch_thread_ = std::thread([this] { dispatch_loop(); });
}

MemCoord::~MemCoord() {
  stop_.store(true);
  {
    std::lock_guard<std::mutex> lk(mu_);
    expiry_cv_.notify_all();
  }
  {
    std::lock_guard<std::mutex> lk(qmu_);
    qcv_.notify_all();
  }
  if (expiry_thread_.joinable()) expiry_thread_.join();
  if (dispatch_thread_.joinable()) dispatch_thread_.join();
}

int64_t MemCoord::now_ms() const {
  return std::chrono::duration_cast<std::chrono::milliseconds>(Clock::now().time_since_epoch()).count() +
         clock_offset_ms_.load();
}

void MemCoord::emit_locked(EventType t, const std::string& key, const std::string& value) {
  if (replaying_) return;  // recovery rebuilds state; nobody is watching yet
  std::lock_guard<std::mutex> lk(qmu_);
  queue_.push_back(WatchEvent{t, key, value, revision_});
  qcv_.notify_one();
}

ErrorCode MemCoord::put_locked(const std::string& key, const std::string& value, LeaseId lease) {
  if (key.empty()) return ErrorCode::INVALID_KEY;
  if (lease != 0 && !leases_.count(lease)) return ErrorCode::ETCD_LEASE_ERROR;
  auto it = kv_.find(key);
  ++revision_;
  if (it == kv_.end()) {
    KeyValue kv{key, value, revision_, revision_, lease};
    kv_.emplace(key, std::move(kv));
  } else {
    if (it->second.lease != lease && it->second.lease != 0) {
      auto l = leases_.find(it->second.lease);
      if (l != leases_.end()) l->second.keys.erase(key);
    }
    it->second.value = value;
    it->second.mod_revision = revision_;
    it->second.lease = lease;
  }
  if (lease != 0) leases_[lease].keys.insert(key);
  if (log_ && !replaying_) {
    wire::Writer w;
    w.u32(OP_PUT);
    w.str(key);
    w.str(value);
    w.i64(lease);
    log_locked(w.data());
  }
  emit_locked(EventType::PUT, key, value);
  return ErrorCode::OK;
}

bool MemCoord::del_locked(const std::string& key) {
  auto it = kv_.find(key);
  if (it == kv_.end()) return false;
  if (it->second.lease != 0) {
    auto l = leases_.find(it->second.lease);
    if (l != leases_.end()) l->second.keys.erase(key);
  }
  const std::string last = std::move(it->second.value);
  kv_.erase(it);
  ++revision_;
  if (log_ && !replaying_) {
    wire::Writer w;
    w.u32(OP_DEL);
    w.str(key);
    log_locked(w.data());
  }
  emit_locked(EventType::DELETE, key, last);
  return true;
}

void MemCoord::expire_locked() {
  const int64_t now = now_ms();
  std::vector<LeaseId> dead;
  for (const auto& [id, l] : leases_)
    if (l.expires_at_ms <= now) dead.push_back(id);
  for (LeaseId id : dead) {
    auto it = leases_.find(id);
    const std::set<std::string> keys = it->second.keys;
    leases_.erase(it);
    if (log_ && !replaying_) {  // expiry is logged like a revoke: a restart must not resurrect the lease's keys
      wire::Writer w;
      w.u32(OP_REVOKE);
      w.i64(id);
      log_locked(w.data());
    }
    for (const auto& k : keys) {
      auto kv = kv_.find(k);
      if (kv != kv_.end() && kv->second.lease == id) {
        kv->second.lease = 0;  // lease is gone; del_locked must not look it up
        del_locked(k);
      }
    }
  }
}

void MemCoord::expiry_loop() {
  std::unique_lock<std::mutex> lk(mu_);
  while (!stop_.load()) {
    expiry_cv_.wait_for(lk, std::chrono::milliseconds(20));
    if (stop_.load()) break;
    expire_locked();
    if (const uint64_t seq = take_seq_locked()) {
      lk.unlock();
      commit(seq, ErrorCode::OK);
      lk.lock();
    }
  }
}

// ---------------------------------------------------------------- durability
void MemCoord::log_locked(const std::string& rec) {
  const uint64_t seq = log_->append(rec);
  if (seq) pending_seq_ = seq;
  else BB_LOG(ERROR) << "coord: metadata log append failed";
}

ErrorCode MemCoord::commit(uint64_t seq, ErrorCode ec) {
  if (!log_ || seq == 0) return ec;
  const ErrorCode sc = log_->sync(seq);
  if (log_->snapshot_due()) {
    std::unique_lock<std::mutex> sl(snap_mu_, std::try_to_lock);
    if (sl.owns_lock() && log_->snapshot_due()) {
      uint64_t gen;
      std::string blob;
      {
        std::lock_guard<std::mutex> lk(mu_);
        gen = log_->rotate();
        blob = snapshot_locked();
      }
      if (log_->install_snapshot(gen, blob) != ErrorCode::OK) BB_LOG(ERROR) << "coord: snapshot " << gen << " failed";
      else BB_LOG(INFO) << "coord: snapshot " << gen << " (" << blob.size() << " bytes)";
    }
  }
  return sc == ErrorCode::OK ? ec : ErrorCode::ETCD_ERROR;
}

std::string MemCoord::snapshot_locked() const {
  wire::Writer w;
  w.i64(revision_);
  w.i64(next_lease_);
  w.u32(static_cast<uint32_t>(leases_.size()));
  for (const auto& [id, l] : leases_) {
    w.i64(id);
    w.i64(l.ttl_ms);
  }
  w.u32(static_cast<uint32_t>(kv_.size()));
  for (const auto& [k, kv] : kv_) {
    w.str(kv.key);
    w.str(kv.value);
    w.i64(kv.create_revision);
    w.i64(kv.mod_revision);
    w.i64(kv.lease);
  }
  return w.take();
}

void MemCoord::load_snapshot(const std::string& blob) {
  wire::Reader r(blob);
  revision_ = r.i64();
  next_lease_ = r.i64();
  const uint32_t nl = r.u32();
  const int64_t now = now_ms();
  for (uint32_t i = 0; i < nl && r.ok(); ++i) {
    const LeaseId id = r.i64();
    const int64_t ttl = r.i64();
    leases_[id] = Lease{ttl, now + ttl, {}};
  }
  const uint32_t nk = r.u32();
  for (uint32_t i = 0; i < nk && r.ok(); ++i) {
    KeyValue kv;
    kv.key = r.str();
    kv.value = r.str();
    kv.create_revision = r.i64();
    kv.mod_revision = r.i64();
    kv.lease = r.i64();
    if (kv.lease) {
      auto l = leases_.find(kv.lease);
      if (l == leases_.end()) continue;  // its lease did not survive: neither does the key
      l->second.keys.insert(kv.key);
    }
    kv_[kv.key] = std::move(kv);
  }
  if (!r.ok()) BB_LOG(ERROR) << "coord: snapshot is truncated";
}

void MemCoord::apply_record(std::string_view rec) {
  wire::Reader r(rec.data(), rec.size());
  switch (r.u32()) {
    case OP_PUT: {
      const std::string k = r.str(), v = r.str();
      const LeaseId l = r.i64();
      if (r.ok()) put_locked(k, v, leases_.count(l) ? l : 0);
      break;
    }
    case OP_DEL: {
      const std::string k = r.str();
      if (r.ok()) del_locked(k);
      break;
    }
    case OP_GRANT: {
      const LeaseId id = r.i64();
      const int64_t ttl = r.i64();
      if (r.ok()) {
        leases_[id] = Lease{ttl, now_ms() + ttl, {}};
        next_lease_ = std::max(next_lease_, id + 1);
      }
      break;
    }
    case OP_REVOKE: {
      const LeaseId id = r.i64();
      auto it = leases_.find(id);
      if (r.ok() && it != leases_.end()) {
        const std::set<std::string> keys = it->second.keys;
        leases_.erase(it);
        for (const auto& k : keys) {
          auto kv = kv_.find(k);
          if (kv != kv_.end() && kv->second.lease == id) {
            kv->second.lease = 0;
            del_locked(k);
          }
        }
      }
      break;
    }
    default: break;
  }
}

ErrorCode MemCoord::open_durable(const std::string& dir, bool fsync, uint64_t snapshot_bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  auto log = std::make_unique<DurableLog>();
  DurableLog::Options o;
  o.dir = dir;
  o.name = "coord";
  o.fsync = fsync;
  o.snapshot_bytes = snapshot_bytes;
  std::string snap;
  std::vector<std::string> recs;
  ErrorCode ec = log->open(o, &snap, [&](std::string_view r) { recs.emplace_back(r); });
  if (ec != ErrorCode::OK) return ec;
  replaying_ = true;
  if (!snap.empty()) load_snapshot(snap);
  for (const auto& r : recs) apply_record(r);
  replaying_ = false;
  recovered_records_ = recs.size();
  // every lease starts a full TTL from now: its holder gets one whole period to find the restarted store
  const int64_t now = now_ms();
  for (auto& [id, l] : leases_) l.expires_at_ms = now + l.ttl_ms;
  log_ = std::move(log);
  if (!snap.empty() || !recs.empty())
    BB_LOG(INFO) << "coord: recovered " << kv_.size() << " keys, " << leases_.size() << " leases at revision " << revision_ << " from " << dir;
  return ErrorCode::OK;
}

void MemCoord::dispatch_loop() {
  while (true) {
    WatchEvent ev;
    {
      std::unique_lock<std::mutex> lk(qmu_);
      qcv_.wait(lk, [this] { return !queue_.empty() || stop_.load(); });
      if (queue_.empty()) {
        if (stop_.load()) return;
        continue;
      }
      ev = std::move(queue_.front());
      queue_.pop_front();
      dispatching_ = true;
    }
    std::vector<std::shared_ptr<Watcher>> targets;
    {
      std::lock_guard<std::mutex> lk(wmu_);
      for (auto& [id, w] : watchers_)
        if (ev.key.compare(0, w->prefix.size(), w->prefix) == 0) {
          ++w->running;
          targets.push_back(w);
        }
    }
    for (auto& w : targets) {
      try {
        w->cb(ev);
      } catch (const std::exception& e) {
        BB_LOG(ERROR) << "watch callback threw: " << e.what();
      }
      std::lock_guard<std::mutex> lk(wmu_);
      --w->running;
      wcv_.notify_all();
    }
    {
      std::lock_guard<std::mutex> lk(qmu_);
      dispatching_ = false;
      if (queue_.empty()) qidle_.notify_all();
    }
  }
}

void MemCoord::flush_events() {
  std::unique_lock<std::mutex> lk(qmu_);
  qidle_.wait_for(lk, std::chrono::seconds(5), [this] { return queue_.empty() && !dispatching_; });
}

void MemCoord::advance_time_ms(int64_t ms) {
  clock_offset_ms_.fetch_add(ms);
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    seq = take_seq_locked();
  }
  commit(seq, ErrorCode::OK);
  flush_events();
}

size_t MemCoord::lease_count() {
  std::lock_guard<std::mutex> lk(mu_);
  return leases_.size();
}
size_t MemCoord::key_count() {
  std::lock_guard<std::mutex> lk(mu_);
  return kv_.size();
}

ErrorCode MemCoord::put(const std::string& key, const std::string& value, LeaseId lease) {
  ErrorCode ec;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ec = put_locked(key, value, lease);
    seq = take_seq_locked();
  }
  return commit(seq, ec);
}

Result<KeyValue> MemCoord::get_kv(const std::string& key) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = kv_.find(key);
  if (it == kv_.end()) return ErrorCode::ETCD_KEY_NOT_FOUND;
  return it->second;
}

ErrorCode MemCoord::del(const std::string& key) {
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    del_locked(key);
    seq = take_seq_locked();
  }
  return commit(seq, ErrorCode::OK);
}

Result<std::vector<KeyValue>> MemCoord::get_with_prefix(const std::string& prefix) {
  std::vector<KeyValue> out;
  std::lock_guard<std::mutex> lk(mu_);
  for (auto it = kv_.lower_bound(prefix); it != kv_.end() && it->first.compare(0, prefix.size(), prefix) == 0; ++it)
    out.push_back(it->second);
  return out;
}

Result<size_t> MemCoord::del_prefix(const std::string& prefix) {
  std::vector<std::string> keys;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = kv_.lower_bound(prefix); it != kv_.end() && it->first.compare(0, prefix.size(), prefix) == 0; ++it)
      keys.push_back(it->first);
    for (const auto& k : keys) del_locked(k);
    seq = take_seq_locked();
  }
  if (commit(seq, ErrorCode::OK) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
  return keys.size();
}

Result<LeaseId> MemCoord::grant_lease(int64_t ttl_sec) {
  if (ttl_sec <= 0) return ErrorCode::INVALID_PARAMETERS;
  LeaseId id;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    id = next_lease_++;
    leases_[id] = Lease{ttl_sec * 1000, now_ms() + ttl_sec * 1000, {}};
    if (log_) {
      wire::Writer w;
      w.u32(OP_GRANT);
      w.i64(id);
      w.i64(ttl_sec * 1000);
      log_locked(w.data());
    }
    seq = take_seq_locked();
  }
  if (commit(seq, ErrorCode::OK) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
  return id;
}

ErrorCode MemCoord::keep_alive(LeaseId lease) {
  ErrorCode ec = ErrorCode::OK;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    auto it = leases_.find(lease);
    if (it == leases_.end()) ec = ErrorCode::ETCD_LEASE_ERROR;
    else it->second.expires_at_ms = now_ms() + it->second.ttl_ms;  // refreshes are not logged: recovery re-arms every lease
    seq = take_seq_locked();
  }
  return commit(seq, ec);
}

ErrorCode MemCoord::revoke_lease(LeaseId lease) {
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = leases_.find(lease);
    if (it == leases_.end()) return ErrorCode::ETCD_LEASE_ERROR;
    it->second.expires_at_ms = 0;
    expire_locked();
    seq = take_seq_locked();
  }
  return commit(seq, ErrorCode::OK);
}

Result<int64_t> MemCoord::lease_remaining_ms(LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = leases_.find(lease);
  if (it == leases_.end()) return ErrorCode::ETCD_LEASE_ERROR;
  return std::max<int64_t>(0, it->second.expires_at_ms - now_ms());
}

Result<bool> MemCoord::put_if_absent(const std::string& key, const std::string& value, LeaseId lease) {
  ErrorCode ec = ErrorCode::OK;
  bool created = false;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    if (!kv_.count(key)) {
      ec = put_locked(key, value, lease);
      created = ec == ErrorCode::OK;
    }
    seq = take_seq_locked();
  }
  ec = commit(seq, ec);
  if (ec != ErrorCode::OK) return ec;
  return created;
}

Result<bool> MemCoord::compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                        LeaseId lease) {
  ErrorCode ec = ErrorCode::OK;
  bool swapped = false;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    auto it = kv_.find(key);
    if (it != kv_.end() && it->second.value == expected) {
      ec = put_locked(key, value, lease);
      swapped = ec == ErrorCode::OK;
    }
    seq = take_seq_locked();
  }
  ec = commit(seq, ec);
  if (ec != ErrorCode::OK) return ec;
  return swapped;
}

Result<bool> MemCoord::compare_and_delete(const std::string& key, const std::string& expected) {
  bool deleted = false;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = kv_.find(key);
    if (it != kv_.end() && it->second.value == expected) deleted = del_locked(key);
    seq = take_seq_locked();
  }
  if (commit(seq, ErrorCode::OK) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
  return deleted;
}

Result<bool> MemCoord::guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                                   const std::string& value) {
  ErrorCode ec = ErrorCode::OK;
  bool done = false;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    auto g = kv_.find(guard_key);
    if (g != kv_.end() && g->second.create_revision == guard_create_revision) {
      ec = put_locked(key, value, 0);
      done = ec == ErrorCode::OK;
    }
    seq = take_seq_locked();
  }
  ec = commit(seq, ec);
  if (ec != ErrorCode::OK) return ec;
  return done;
}

Result<bool> MemCoord::guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) {
  bool done = false;
  uint64_t seq;
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
    auto g = kv_.find(guard_key);
    if (g != kv_.end() && g->second.create_revision == guard_create_revision) {
      del_locked(key);
      done = true;
    }
    seq = take_seq_locked();
  }
  if (commit(seq, ErrorCode::OK) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
  return done;
}

Result<int64_t> MemCoord::watch_prefix(const std::string& prefix, WatchCallback cb) {
  if (!cb) return ErrorCode::ETCD_WATCH_ERROR;
  std::lock_guard<std::mutex> lk(wmu_);
  const int64_t id = next_watch_++;
  watchers_[id] = std::make_shared<Watcher>(Watcher{prefix, std::move(cb)});
  return id;
}

ErrorCode MemCoord::unwatch(int64_t id) {
  std::unique_lock<std::mutex> lk(wmu_);
  auto it = watchers_.find(id);
  if (it == watchers_.end()) return ErrorCode::ETCD_WATCH_ERROR;
  std::shared_ptr<Watcher> w = it->second;
  watchers_.erase(it);
  // barrier: an invocation already handed to the dispatch thread finishes before we return (unless we ARE that thread)
  if (std::this_thread::get_id() != dispatch_thread_.get_id()) wcv_.wait(lk, [&] { return w->running == 0; });
  return ErrorCode::OK;
}

int64_t MemCoord::revision() {
  std::lock_guard<std::mutex> lk(mu_);
  return revision_;
}

// ================================================================ wire protocol
namespace {
enum Method : uint32_t {
  M_PUT = 1, M_GET, M_DEL, M_PREFIX, M_DEL_PREFIX, M_GRANT, M_KEEPALIVE, M_REVOKE, M_REMAINING,
  M_PUT_IF_ABSENT, M_CAS, M_CAD, M_WATCH, M_UNWATCH, M_REVISION, M_GUARDED_PUT, M_GUARDED_DEL,
};
constexpr uint32_t kTopicWatch = 1;

void put_kv(wire::Writer& w, const KeyValue& kv) {
  w.str(kv.key);
  w.str(kv.value);
  w.i64(kv.create_revision);
  w.i64(kv.mod_revision);
  w.i64(kv.lease);
}
KeyValue get_kv_wire(wire::Reader& r) {
  KeyValue kv;
  kv.key = r.str();
  kv.value = r.str();
  kv.create_revision = r.i64();
  kv.mod_revision = r.i64();
  kv.lease = r.i64();
  return kv;
}
template <typename T, typename F>
std::string reply(const Result<T>& res, F&& enc) {
  wire::Writer w;
  w.ec(res.error());
  if (res.ok()) enc(w, res.value());
  return w.take();
}
std::string reply_ec(ErrorCode ec) {
  wire::Writer w;
  w.ec(ec);
  return w.take();
}
}  // namespace

struct CoordServer::ConnState {
  std::mutex mu;
  std::vector<int64_t> watches;
};

CoordServer::CoordServer(std::shared_ptr<MemCoord> store) : store_(store ? std::move(store) : std::make_shared<MemCoord>()) {
  auto st = store_;
  // read-only members may look and listen, nothing else (no writes, no leases, no elections)
  rpc_.allow_read_only({M_GET, M_PREFIX, M_REMAINING, M_REVISION, M_WATCH, M_UNWATCH});
  rpc_.register_method(M_PUT, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string k = r.str(), v = r.str();
    LeaseId l = r.i64();
    return reply_ec(r.ok() ? st->put(k, v, l) : ErrorCode::INVALID_PARAMETERS);
  });
  rpc_.register_method(M_GET, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply(st->get_kv(r.str()), [](wire::Writer& w, const KeyValue& kv) { put_kv(w, kv); });
  });
  rpc_.register_method(M_DEL, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply_ec(st->del(r.str()));
  });
  rpc_.register_method(M_PREFIX, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply(st->get_with_prefix(r.str()), [](wire::Writer& w, const std::vector<KeyValue>& v) {
      w.u32(static_cast<uint32_t>(v.size()));
      for (const auto& kv : v) put_kv(w, kv);
    });
  });
  rpc_.register_method(M_DEL_PREFIX, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply(st->del_prefix(r.str()), [](wire::Writer& w, size_t n) { w.u64(n); });
  });
  rpc_.register_method(M_GRANT, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply(st->grant_lease(r.i64()), [](wire::Writer& w, LeaseId l) { w.i64(l); });
  });
  rpc_.register_method(M_KEEPALIVE, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply_ec(st->keep_alive(r.i64()));
  });
  rpc_.register_method(M_REVOKE, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply_ec(st->revoke_lease(r.i64()));
  });
  rpc_.register_method(M_REMAINING, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply(st->lease_remaining_ms(r.i64()), [](wire::Writer& w, int64_t v) { w.i64(v); });
  });
  rpc_.register_method(M_PUT_IF_ABSENT, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string k = r.str(), v = r.str();
    LeaseId l = r.i64();
    return reply(st->put_if_absent(k, v, l), [](wire::Writer& w, bool b) { w.boolean(b); });
  });
  rpc_.register_method(M_CAS, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string k = r.str(), e = r.str(), v = r.str();
    LeaseId l = r.i64();
    return reply(st->compare_and_swap(k, e, v, l), [](wire::Writer& w, bool b) { w.boolean(b); });
  });
  rpc_.register_method(M_CAD, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string k = r.str(), e = r.str();
    return reply(st->compare_and_delete(k, e), [](wire::Writer& w, bool b) { w.boolean(b); });
  });
  rpc_.register_method(M_GUARDED_PUT, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string g = r.str();
    const int64_t rev = r.i64();
    std::string k = r.str(), v = r.str();
    if (!r.ok()) return reply_ec(ErrorCode::INVALID_PARAMETERS);
    return reply(st->guarded_put(g, rev, k, v), [](wire::Writer& w, bool b) { w.boolean(b); });
  });
  rpc_.register_method(M_GUARDED_DEL, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    std::string g = r.str();
    const int64_t rev = r.i64();
    std::string k = r.str();
    if (!r.ok()) return reply_ec(ErrorCode::INVALID_PARAMETERS);
    return reply(st->guarded_del(g, rev, k), [](wire::Writer& w, bool b) { w.boolean(b); });
  });
  rpc_.register_method(M_REVISION, [st](const net::ConnPtr&, const std::string&) {
    wire::Writer w;
    w.ec(ErrorCode::OK);
    w.i64(st->revision());
    return w.take();
  });
  rpc_.register_method(M_WATCH, [st](const net::ConnPtr& c, const std::string& q) {
    wire::Reader r(q);
    const std::string prefix = r.str();
    if (!c->user) c->user = std::make_shared<ConnState>();
    auto cs = std::static_pointer_cast<ConnState>(c->user);
    std::weak_ptr<net::Connection> weak = c;
    auto id_holder = std::make_shared<std::atomic<int64_t>>(0);
    auto res = st->watch_prefix(prefix, [weak, id_holder](const WatchEvent& ev) {
      auto conn = weak.lock();
      if (!conn) return;
      wire::Writer w;
      w.i64(id_holder->load());
      w.u32(static_cast<uint32_t>(ev.type));
      w.str(ev.key);
      w.str(ev.value);
      w.i64(ev.revision);
      net::RpcServer::push(conn, kTopicWatch, w.data());
    });
    if (res.ok()) {
      id_holder->store(res.value());
      std::lock_guard<std::mutex> lk(cs->mu);
      cs->watches.push_back(res.value());
    }
    return reply(res, [](wire::Writer& w, int64_t id) { w.i64(id); });
  });
  rpc_.register_method(M_UNWATCH, [st](const net::ConnPtr&, const std::string& q) {
    wire::Reader r(q);
    return reply_ec(st->unwatch(r.i64()));
  });
  rpc_.set_close_hook([st](const net::ConnPtr& c) {
    if (!c->user) return;
    auto cs = std::static_pointer_cast<ConnState>(c->user);
    std::lock_guard<std::mutex> lk(cs->mu);
    for (int64_t id : cs->watches) st->unwatch(id);
    cs->watches.clear();
  });
}

CoordServer::~CoordServer() { stop(); }
ErrorCode CoordServer::start(const std::string& host, uint16_t port) { return rpc_.start(host, port, 2); }
void CoordServer::stop() { rpc_.stop(); }

// ================================================================ RemoteCoord
// Connection loss is survivable (the daemon may be restarted with --data-dir and come back with the same keys, revisions
// and lease ids): request/response calls reconnect and retry once; the watch channel is re-established by a monitor
// thread which re-registers every watch and then *re-lists* its prefix, delivering PUTs for what exists now and DELETEs
// for keys that were seen before and are gone -- so a watcher converges even though events were missed while the
// store was away.  Handlers must be idempotent (the Keystone's and the workers' are).
namespace {
thread_local const RemoteCoord* tl_delivering = nullptr;  // set while this thread runs a watch callback of that store
}

RemoteCoord::~RemoteCoord() { close(); }

bool RemoteCoord::connect_any(net::RpcClient& c, int timeout_ms) {
  for (const auto& [h, p] : endpoints_)
    if (c.connect(h, p, timeout_ms) == ErrorCode::OK) return true;
  return false;
}

ErrorCode RemoteCoord::connect(const std::string& endpoints, int timeout_ms) {
  std::stringstream ss(endpoints);
  std::string ep;
  endpoints_.clear();
  while (std::getline(ss, ep, ',')) {
    while (!ep.empty() && ep.front() == ' ') ep.erase(ep.begin());
    const std::string pfx = "tcp://";
    if (ep.compare(0, pfx.size(), pfx) == 0) ep = ep.substr(pfx.size());
    auto hp = split_host_port(ep);
    if (hp) endpoints_.emplace_back(hp->first, static_cast<uint16_t>(hp->second));
  }
  closing_.store(false);
  std::lock_guard<std::mutex> lk(conn_mu_);
  return connect_any(rpc_, timeout_ms) ? ErrorCode::OK : ErrorCode::CONNECTION_FAILED;
}

void RemoteCoord::close() {
  closing_.store(true);
  if (monitor_.joinable()) monitor_.join();
  rpc_.close();
  watch_rpc_.close();
}

Result<std::string> RemoteCoord::call(uint32_t method, const std::string& req) {
  for (int attempt = 0; attempt < 2; ++attempt) {
    const uint64_t gen = conn_gen_.load();
    auto r = rpc_.call(method, req, 10000);
    if (r.ok()) return r;
    if (closing_.load() || (r.error() != ErrorCode::RPC_FAILED && r.error() != ErrorCode::CLIENT_DISCONNECTED)) break;
    // the store went away: look for it again (same endpoint list) and retry the request once
    std::lock_guard<std::mutex> lk(conn_mu_);
    if (conn_gen_.load() != gen) continue;  // another caller already reconnected
    if (!connect_any(rpc_, 1000)) break;
    conn_gen_.fetch_add(1);
    reconnects_.fetch_add(1);
  }
  return ErrorCode::ETCD_ERROR;
}

#define BB_COORD_CALL(method, writer)                \
  auto _resp = call(method, (writer).data());        \
  if (!_resp.ok()) return _resp.error();             \
  wire::Reader rd(_resp.value());                    \
  const ErrorCode _ec = rd.ec();                     \
  if (!rd.ok()) return ErrorCode::ETCD_ERROR;

ErrorCode RemoteCoord::put(const std::string& key, const std::string& value, LeaseId lease) {
  wire::Writer w;
  w.str(key);
  w.str(value);
  w.i64(lease);
  BB_COORD_CALL(M_PUT, w);
  return _ec;
}
Result<KeyValue> RemoteCoord::get_kv(const std::string& key) {
  wire::Writer w;
  w.str(key);
  BB_COORD_CALL(M_GET, w);
  if (_ec != ErrorCode::OK) return _ec;
  return get_kv_wire(rd);
}
ErrorCode RemoteCoord::del(const std::string& key) {
  wire::Writer w;
  w.str(key);
  BB_COORD_CALL(M_DEL, w);
  return _ec;
}
Result<std::vector<KeyValue>> RemoteCoord::get_with_prefix(const std::string& prefix) {
  wire::Writer w;
  w.str(prefix);
  BB_COORD_CALL(M_PREFIX, w);
  if (_ec != ErrorCode::OK) return _ec;
  const uint32_t n = rd.count(8);
  std::vector<KeyValue> v;
  v.reserve(n);
  for (uint32_t i = 0; i < n; ++i) v.push_back(get_kv_wire(rd));
  if (!rd.ok()) return ErrorCode::ETCD_ERROR;
  return v;
}
Result<size_t> RemoteCoord::del_prefix(const std::string& prefix) {
  wire::Writer w;
  w.str(prefix);
  BB_COORD_CALL(M_DEL_PREFIX, w);
  if (_ec != ErrorCode::OK) return _ec;
  return static_cast<size_t>(rd.u64());
}
Result<LeaseId> RemoteCoord::grant_lease(int64_t ttl_sec) {
  wire::Writer w;
  w.i64(ttl_sec);
  BB_COORD_CALL(M_GRANT, w);
  if (_ec != ErrorCode::OK) return _ec;
  return static_cast<LeaseId>(rd.i64());
}
ErrorCode RemoteCoord::keep_alive(LeaseId lease) {
  wire::Writer w;
  w.i64(lease);
  BB_COORD_CALL(M_KEEPALIVE, w);
  return _ec;
}
ErrorCode RemoteCoord::revoke_lease(LeaseId lease) {
  wire::Writer w;
  w.i64(lease);
  BB_COORD_CALL(M_REVOKE, w);
  return _ec;
}
Result<int64_t> RemoteCoord::lease_remaining_ms(LeaseId lease) {
  wire::Writer w;
  w.i64(lease);
  BB_COORD_CALL(M_REMAINING, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.i64();
}
Result<bool> RemoteCoord::put_if_absent(const std::string& key, const std::string& value, LeaseId lease) {
  wire::Writer w;
  w.str(key);
  w.str(value);
  w.i64(lease);
  BB_COORD_CALL(M_PUT_IF_ABSENT, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.boolean();
}
Result<bool> RemoteCoord::compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                           LeaseId lease) {
  wire::Writer w;
  w.str(key);
  w.str(expected);
  w.str(value);
  w.i64(lease);
  BB_COORD_CALL(M_CAS, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.boolean();
}
Result<bool> RemoteCoord::compare_and_delete(const std::string& key, const std::string& expected) {
  wire::Writer w;
  w.str(key);
  w.str(expected);
  BB_COORD_CALL(M_CAD, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.boolean();
}
int64_t RemoteCoord::revision() {
  auto resp = call(M_REVISION, "");
  if (!resp.ok()) return -1;
  wire::Reader rd(resp.value());
  rd.ec();
  return rd.i64();
}

Result<bool> RemoteCoord::guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                                      const std::string& value) {
  wire::Writer w;
  w.str(guard_key);
  w.i64(guard_create_revision);
  w.str(key);
  w.str(value);
  BB_COORD_CALL(M_GUARDED_PUT, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.boolean();
}
Result<bool> RemoteCoord::guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) {
  wire::Writer w;
  w.str(guard_key);
  w.i64(guard_create_revision);
  w.str(key);
  BB_COORD_CALL(M_GUARDED_DEL, w);
  if (_ec != ErrorCode::OK) return _ec;
  return rd.boolean();
}

// Runs `cb` for one event with the bookkeeping unwatch()'s barrier relies on.
void RemoteCoord::deliver(int64_t local_id, const WatchEvent& ev) {
  WatchCallback f;
  {
    std::lock_guard<std::mutex> lk(wmu_);
    auto it = watches_.find(local_id);
    if (it == watches_.end()) return;
    f = it->second.cb;
    if (ev.type == EventType::PUT) it->second.known.insert(ev.key);
    else it->second.known.erase(ev.key);
    ++running_[local_id];
  }
  const RemoteCoord* prev = tl_delivering;
  tl_delivering = this;
  try {
    f(ev);
  } catch (const std::exception& e) {
    BB_LOG(ERROR) << "watch callback threw: " << e.what();
  }
  tl_delivering = prev;
  std::lock_guard<std::mutex> lk(wmu_);
  if (--running_[local_id] == 0) running_.erase(local_id);
  wcv_.notify_all();
}

bool RemoteCoord::open_watch_channel() {
  if (!connect_any(watch_rpc_, 1000)) return false;
  watch_rpc_.enable_push([this](uint32_t topic, const std::string& payload) {
    if (topic != kTopicWatch) return;
    wire::Reader r(payload);
    const int64_t server_id = r.i64();
    WatchEvent ev;
    ev.type = static_cast<EventType>(r.u32());
    ev.key = r.str();
    ev.value = r.str();
    ev.revision = r.i64();
    if (!r.ok()) return;
    int64_t local = 0;
    {
      std::lock_guard<std::mutex> lk(wmu_);
      auto it = by_server_.find(server_id);
      if (it == by_server_.end()) {
        pending_[server_id].push_back(ev);  // event raced ahead of the watch response
        return;
      }
      local = it->second;
    }
    deliver(local, ev);
  });
  return true;
}

Result<int64_t> RemoteCoord::server_watch(const std::string& prefix) {
  wire::Writer w;
  w.str(prefix);
  auto resp = watch_rpc_.call(M_WATCH, w.data(), 10000);
  if (!resp.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  wire::Reader rd(resp.value());
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  return rd.i64();
}

Result<int64_t> RemoteCoord::watch_prefix(const std::string& prefix, WatchCallback cb) {
  std::lock_guard<std::mutex> cl(watch_conn_mu_);
  if (!watch_connected_) {
    if (!open_watch_channel()) return ErrorCode::ETCD_WATCH_ERROR;
    watch_connected_ = true;
    if (!monitor_.joinable()) monitor_ = std::thread([this] { monitor_loop(); });
  }
  auto sid = server_watch(prefix);
  if (!sid.ok()) return sid.error();
  int64_t local;
  std::vector<WatchEvent> early;
  {
    std::lock_guard<std::mutex> lk(wmu_);
    local = next_local_watch_++;
    WatchReg reg;
    reg.prefix = prefix;
    reg.cb = std::move(cb);
    reg.server_id = sid.value();
    watches_[local] = std::move(reg);
    by_server_[sid.value()] = local;
    auto it = pending_.find(sid.value());
    if (it != pending_.end()) {
      early = std::move(it->second);
      pending_.erase(it);
    }
  }
  for (const auto& ev : early) deliver(local, ev);
  return local;
}

ErrorCode RemoteCoord::unwatch(int64_t id) {
  int64_t server_id = 0;
  {
    std::unique_lock<std::mutex> lk(wmu_);
    auto it = watches_.find(id);
    if (it == watches_.end()) return ErrorCode::ETCD_WATCH_ERROR;
    server_id = it->second.server_id;
    by_server_.erase(server_id);
    pending_.erase(server_id);
    watches_.erase(it);
    // barrier: a callback already running on another thread finishes before we return (unless we ARE inside one)
    if (tl_delivering != this) wcv_.wait(lk, [&] { return running_.find(id) == running_.end(); });
  }
  if (tl_delivering == this) return ErrorCode::OK;  // server side is dropped with the connection at the latest
  std::lock_guard<std::mutex> cl(watch_conn_mu_);
  if (!watch_connected_ || !watch_rpc_.healthy()) return ErrorCode::OK;
  wire::Writer w;
  w.i64(server_id);
  auto resp = watch_rpc_.call(M_UNWATCH, w.data(), 10000);
  if (!resp.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  wire::Reader rd(resp.value());
  return rd.ec();
}

// Re-establishes the push channel after the store went away, then replays the difference to every watcher.
void RemoteCoord::monitor_loop() {
  while (!closing_.load()) {
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    if (closing_.load()) break;
    {
      std::lock_guard<std::mutex> cl(watch_conn_mu_);
      if (!watch_connected_ || watch_rpc_.healthy()) continue;
      if (!open_watch_channel()) continue;  // still down: try again on the next tick
      reconnects_.fetch_add(1);
      std::vector<std::pair<int64_t, std::string>> regs;
      {
        std::lock_guard<std::mutex> lk(wmu_);
        by_server_.clear();
        pending_.clear();
        for (const auto& [id, r] : watches_) regs.emplace_back(id, r.prefix);
      }
      bool ok = true;
      for (const auto& [local, prefix] : regs) {
        auto sid = server_watch(prefix);
        if (!sid.ok()) {
          ok = false;
          break;
        }
        std::lock_guard<std::mutex> lk(wmu_);
        auto it = watches_.find(local);
        if (it == watches_.end()) continue;
        it->second.server_id = sid.value();
        by_server_[sid.value()] = local;
      }
      if (!ok) {
        watch_rpc_.close();
        continue;
      }
      BB_LOG(INFO) << "coord client: watch channel re-established (" << regs.size() << " watches), re-listing";
    }
    // ---- resync outside the connection lock: callbacks may call back into this store
    std::vector<std::pair<int64_t, std::string>> regs;
    {
      std::lock_guard<std::mutex> lk(wmu_);
      for (const auto& [id, r] : watches_) regs.emplace_back(id, r.prefix);
    }
    for (const auto& [local, prefix] : regs) {
      auto now = get_with_prefix(prefix);
      if (!now.ok()) continue;
      std::set<std::string> present;
      for (const auto& kv : now.value()) present.insert(kv.key);
      std::vector<std::string> gone;
      {
        std::lock_guard<std::mutex> lk(wmu_);
        auto it = watches_.find(local);
        if (it == watches_.end()) continue;
        for (const auto& k : it->second.known)
          if (!present.count(k)) gone.push_back(k);
      }
      for (const auto& k : gone) deliver(local, WatchEvent{EventType::DELETE, k, "", 0});
      for (const auto& kv : now.value()) deliver(local, WatchEvent{EventType::PUT, kv.key, kv.value, kv.mod_revision});
    }
  }
}

// ================================================================ shared in-proc stores
namespace {
std::mutex g_shared_mu;
std::map<std::string, std::shared_ptr<MemCoord>> g_shared;
}  // namespace

std::shared_ptr<MemCoord> shared_mem_coord(const std::string& name) {
  std::lock_guard<std::mutex> lk(g_shared_mu);
  auto& s = g_shared[name];
  if (!s) s = std::make_shared<MemCoord>();
  return s;
}
void drop_shared_mem_coord(const std::string& name) {
  std::lock_guard<std::mutex> lk(g_shared_mu);
  g_shared.erase(name);
}

// ================================================================ CoordService
CoordService::CoordService(const std::string& endpoints) : endpoints_(endpoints) {}
CoordService::CoordService(std::shared_ptr<CoordStore> store) : store_(std::move(store)), connected_(store_ != nullptr) {}

CoordService::~CoordService() {
  if (store_)
    for (int64_t id : watch_ids_) store_->unwatch(id);
}

ErrorCode CoordService::connect() {
  if (connected_) return ErrorCode::OK;
  const std::string mem = "mem://";
  if (endpoints_.empty() || endpoints_ == mem) {
    store_ = std::make_shared<MemCoord>();
  } else if (endpoints_.compare(0, mem.size(), mem) == 0) {
    store_ = shared_mem_coord(endpoints_.substr(mem.size()));
  } else if (endpoints_.compare(0, 7, "etcd://") == 0) {
    // a real etcd v3 cluster through its JSON gateway (coord/etcd_coord.h); same key schema, same CoordStore semantics
    auto ec = std::make_shared<EtcdCoord>();
    if (ec->connect(endpoints_) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
    store_ = ec;
  } else {
    auto rc = std::make_shared<RemoteCoord>();
    if (rc->connect(endpoints_) != ErrorCode::OK) return ErrorCode::ETCD_ERROR;
    store_ = rc;
  }
  connected_ = true;
  return ErrorCode::OK;
}

ErrorCode CoordService::get(const std::string& key, std::string& value) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  auto r = store_->get(key);
  if (!r.ok()) return r.error();
  value = r.value();
  return ErrorCode::OK;
}
ErrorCode CoordService::put(const std::string& key, const std::string& value) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  return store_->put(key, value, 0);
}
ErrorCode CoordService::put_with_ttl(const std::string& key, const std::string& value, int64_t ttl_sec) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  LeaseId lease = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = ttl_leases_.find(key);
    if (it != ttl_leases_.end()) lease = it->second;
  }
  if (lease != 0 && store_->keep_alive(lease) == ErrorCode::OK) {
    if (store_->put(key, value, lease) == ErrorCode::OK) return ErrorCode::OK;
  }
  auto g = store_->grant_lease(ttl_sec);
  if (!g.ok()) return ErrorCode::ETCD_LEASE_ERROR;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ttl_leases_[key] = g.value();
  }
  return store_->put(key, value, g.value());
}
ErrorCode CoordService::del(const std::string& key) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  LeaseId lease = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = ttl_leases_.find(key);
    if (it != ttl_leases_.end()) {
      lease = it->second;
      ttl_leases_.erase(it);
    }
  }
  if (lease) store_->revoke_lease(lease);
  return store_->del(key);
}
ErrorCode CoordService::get_with_prefix(const std::string& prefix, std::vector<std::string>& keys, std::vector<std::string>& values) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  auto r = store_->get_with_prefix(prefix);
  if (!r.ok()) return r.error();
  keys.clear();
  values.clear();
  for (auto& kv : r.value()) {
    keys.push_back(kv.key);
    values.push_back(kv.value);
  }
  return ErrorCode::OK;
}
ErrorCode CoordService::grant_lease(int64_t ttl_sec, LeaseId& lease) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  auto g = store_->grant_lease(ttl_sec);
  if (!g.ok()) return g.error();
  lease = g.value();
  return ErrorCode::OK;
}
ErrorCode CoordService::put_with_lease(const std::string& key, const std::string& value, LeaseId lease) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  return store_->put(key, value, lease);
}
ErrorCode CoordService::keep_alive(LeaseId lease) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  return store_->keep_alive(lease);
}
ErrorCode CoordService::revoke_lease(LeaseId lease) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  return store_->revoke_lease(lease);
}
ErrorCode CoordService::watch_prefix(const std::string& prefix, WatchCb cb, int64_t* watch_id) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  auto r = store_->watch_prefix(prefix, [cb](const WatchEvent& ev) { cb(ev.key, ev.value, ev.type == EventType::DELETE); });
  if (!r.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  if (watch_id) *watch_id = r.value();
  std::lock_guard<std::mutex> lk(mu_);
  watch_ids_.push_back(r.value());
  return ErrorCode::OK;
}
ErrorCode CoordService::unwatch(int64_t watch_id) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = std::find(watch_ids_.begin(), watch_ids_.end(), watch_id);
    if (it == watch_ids_.end()) return ErrorCode::ETCD_WATCH_ERROR;
    watch_ids_.erase(it);
  }
  return store_->unwatch(watch_id);  // barrier (see coord.h)
}
ErrorCode CoordService::watch_key(const std::string& key, WatchCb cb) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  auto r = store_->watch_prefix(key, [cb, key](const WatchEvent& ev) {
    if (ev.key == key) cb(ev.key, ev.value, ev.type == EventType::DELETE);
  });
  if (!r.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  std::lock_guard<std::mutex> lk(mu_);
  key_watches_[key] = r.value();
  return ErrorCode::OK;
}
ErrorCode CoordService::unwatch_key(const std::string& key) {
  int64_t id = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = key_watches_.find(key);
    if (it == key_watches_.end()) return ErrorCode::ETCD_WATCH_ERROR;
    id = it->second;
    key_watches_.erase(it);
  }
  return store_->unwatch(id);
}

ErrorCode CoordService::register_service(const std::string& name, const std::string& id, const std::string& address, int64_t ttl_sec) {
  ErrorCode ec = put_with_ttl("/blackbird/services/" + name + "/" + id, address, ttl_sec);
  return ec == ErrorCode::OK ? ec : ErrorCode::SERVICE_REGISTRATION_FAILED;
}
ErrorCode CoordService::discover_service(const std::string& name, std::vector<std::string>& addresses) {
  std::vector<std::string> keys;
  return get_with_prefix("/blackbird/services/" + name + "/", keys, addresses);
}
ErrorCode CoordService::unregister_service(const std::string& name, const std::string& id) {
  return del("/blackbird/services/" + name + "/" + id);
}

ErrorCode CoordService::campaign_leader(const std::string& election, const std::string& candidate, int64_t ttl_sec, bool& is_leader) {
  is_leader = false;
  if (!connected_) return ErrorCode::ETCD_ERROR;
  const std::string key = "/blackbird/elections/" + election + "/leader";
  // already the leader? refresh
  auto cur = store_->get_kv(key);
  if (cur.ok() && cur.value().value == candidate) {
    LeaseId lease = cur.value().lease;
    if (lease && store_->keep_alive(lease) == ErrorCode::OK) {
      std::lock_guard<std::mutex> lk(mu_);
      election_leases_[election] = lease;
      election_terms_[election] = cur.value().create_revision;
      is_leader = true;
      return ErrorCode::OK;
    }
  }
  auto g = store_->grant_lease(ttl_sec);
  if (!g.ok()) return ErrorCode::LEADER_ELECTION_FAILED;
  auto won = store_->put_if_absent(key, candidate, g.value());
  if (!won.ok()) {
    store_->revoke_lease(g.value());
    return ErrorCode::LEADER_ELECTION_FAILED;
  }
  if (won.value()) {
    auto mine = store_->get_kv(key);  // the term = create revision of the key we just created
    if (!mine.ok() || mine.value().value != candidate) {
      store_->revoke_lease(g.value());
      return ErrorCode::LEADER_ELECTION_FAILED;
    }
    std::lock_guard<std::mutex> lk(mu_);
    election_leases_[election] = g.value();
    election_terms_[election] = mine.value().create_revision;
    is_leader = true;
  } else {
    store_->revoke_lease(g.value());
  }
  return ErrorCode::OK;
}

int64_t CoordService::leader_term(const std::string& election) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = election_terms_.find(election);
  return it == election_terms_.end() ? 0 : it->second;
}

ErrorCode CoordService::fenced_put(const std::string& election, const std::string& key, const std::string& value) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  const int64_t term = leader_term(election);
  if (term == 0) return ErrorCode::NOT_LEADER;
  auto r = store_->guarded_put("/blackbird/elections/" + election + "/leader", term, key, value);
  if (!r.ok()) return r.error();
  return r.value() ? ErrorCode::OK : ErrorCode::NOT_LEADER;
}

ErrorCode CoordService::fenced_del(const std::string& election, const std::string& key) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  const int64_t term = leader_term(election);
  if (term == 0) return ErrorCode::NOT_LEADER;
  auto r = store_->guarded_del("/blackbird/elections/" + election + "/leader", term, key);
  if (!r.ok()) return r.error();
  return r.value() ? ErrorCode::OK : ErrorCode::NOT_LEADER;
}

ErrorCode CoordService::get_leader(const std::string& election, std::string& leader) {
  return get("/blackbird/elections/" + election + "/leader", leader);
}

ErrorCode CoordService::resign_leader(const std::string& election, const std::string& candidate) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  const std::string key = "/blackbird/elections/" + election + "/leader";
  auto r = store_->compare_and_delete(key, candidate);
  LeaseId lease = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = election_leases_.find(election);
    if (it != election_leases_.end()) {
      lease = it->second;
      election_leases_.erase(it);
    }
    election_terms_.erase(election);
  }
  if (lease) store_->revoke_lease(lease);
  if (!r.ok()) return r.error();
  return r.value() ? ErrorCode::OK : ErrorCode::NOT_LEADER;
}

ErrorCode CoordService::refresh_leadership(const std::string& election, const std::string& candidate) {
  if (!connected_) return ErrorCode::ETCD_ERROR;
  LeaseId lease = 0;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = election_leases_.find(election);
    if (it != election_leases_.end()) lease = it->second;
  }
  auto lost = [&] {
    std::lock_guard<std::mutex> lk(mu_);
    election_terms_.erase(election);
    return ErrorCode::NOT_LEADER;
  };
  if (!lease) return lost();
  const ErrorCode ka = store_->keep_alive(lease);
  if (ka == ErrorCode::ETCD_ERROR) return ErrorCode::ETCD_ERROR;  // store unreachable: unknown, not (yet) lost
  if (ka != ErrorCode::OK) return lost();
  auto cur = store_->get_kv("/blackbird/elections/" + election + "/leader");
  if (!cur.ok() && cur.error() == ErrorCode::ETCD_ERROR) return ErrorCode::ETCD_ERROR;
  if (!cur.ok() || cur.value().value != candidate || cur.value().create_revision != leader_term(election)) return lost();
  return ErrorCode::OK;
}

}  // namespace bb::coord
