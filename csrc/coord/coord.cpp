#include "coord/coord.h"

#include <algorithm>
#include <chrono>
#include <sstream>

#include "common/log.h"
#include "rpc/wire.h"

namespace bb::coord {

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
  }
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
  {
    std::lock_guard<std::mutex> lk(mu_);
    expire_locked();
  }
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
  std::lock_guard<std::mutex> lk(mu_);
  return put_locked(key, value, lease);
}

Result<KeyValue> MemCoord::get_kv(const std::string& key) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = kv_.find(key);
  if (it == kv_.end()) return ErrorCode::ETCD_KEY_NOT_FOUND;
  return it->second;
}

ErrorCode MemCoord::del(const std::string& key) {
  std::lock_guard<std::mutex> lk(mu_);
  del_locked(key);
  return ErrorCode::OK;
}

Result<std::vector<KeyValue>> MemCoord::get_with_prefix(const std::string& prefix) {
  std::vector<KeyValue> out;
  std::lock_guard<std::mutex> lk(mu_);
  for (auto it = kv_.lower_bound(prefix); it != kv_.end() && it->first.compare(0, prefix.size(), prefix) == 0; ++it)
    out.push_back(it->second);
  return out;
}

Result<size_t> MemCoord::del_prefix(const std::string& prefix) {
  std::lock_guard<std::mutex> lk(mu_);
  std::vector<std::string> keys;
  for (auto it = kv_.lower_bound(prefix); it != kv_.end() && it->first.compare(0, prefix.size(), prefix) == 0; ++it)
    keys.push_back(it->first);
  for (const auto& k : keys) del_locked(k);
  return keys.size();
}

Result<LeaseId> MemCoord::grant_lease(int64_t ttl_sec) {
  if (ttl_sec <= 0) return ErrorCode::INVALID_PARAMETERS;
  std::lock_guard<std::mutex> lk(mu_);
  const LeaseId id = next_lease_++;
  leases_[id] = Lease{ttl_sec * 1000, now_ms() + ttl_sec * 1000, {}};
  return id;
}

ErrorCode MemCoord::keep_alive(LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  expire_locked();
  auto it = leases_.find(lease);
  if (it == leases_.end()) return ErrorCode::ETCD_LEASE_ERROR;
  it->second.expires_at_ms = now_ms() + it->second.ttl_ms;
  return ErrorCode::OK;
}

ErrorCode MemCoord::revoke_lease(LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = leases_.find(lease);
  if (it == leases_.end()) return ErrorCode::ETCD_LEASE_ERROR;
  it->second.expires_at_ms = 0;
  expire_locked();
  return ErrorCode::OK;
}

Result<int64_t> MemCoord::lease_remaining_ms(LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = leases_.find(lease);
  if (it == leases_.end()) return ErrorCode::ETCD_LEASE_ERROR;
  return std::max<int64_t>(0, it->second.expires_at_ms - now_ms());
}

Result<bool> MemCoord::put_if_absent(const std::string& key, const std::string& value, LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  expire_locked();
  if (kv_.count(key)) return false;
  ErrorCode ec = put_locked(key, value, lease);
  if (ec != ErrorCode::OK) return ec;
  return true;
}

Result<bool> MemCoord::compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                        LeaseId lease) {
  std::lock_guard<std::mutex> lk(mu_);
  expire_locked();
  auto it = kv_.find(key);
  if (it == kv_.end() || it->second.value != expected) return false;
  ErrorCode ec = put_locked(key, value, lease);
  if (ec != ErrorCode::OK) return ec;
  return true;
}

Result<bool> MemCoord::compare_and_delete(const std::string& key, const std::string& expected) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = kv_.find(key);
  if (it == kv_.end() || it->second.value != expected) return false;
  del_locked(key);
  return true;
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
  M_PUT_IF_ABSENT, M_CAS, M_CAD, M_WATCH, M_UNWATCH, M_REVISION,
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
RemoteCoord::~RemoteCoord() { close(); }

ErrorCode RemoteCoord::connect(const std::string& endpoints, int timeout_ms) {
  std::stringstream ss(endpoints);
  std::string ep;
  while (std::getline(ss, ep, ',')) {
    while (!ep.empty() && ep.front() == ' ') ep.erase(ep.begin());
    const std::string pfx = "tcp://";
    if (ep.compare(0, pfx.size(), pfx) == 0) ep = ep.substr(pfx.size());
    auto hp = split_host_port(ep);
    if (!hp) continue;
    if (rpc_.connect(hp->first, static_cast<uint16_t>(hp->second), timeout_ms) == ErrorCode::OK) {
      host_ = hp->first;
      port_ = static_cast<uint16_t>(hp->second);
      return ErrorCode::OK;
    }
  }
  return ErrorCode::CONNECTION_FAILED;
}

void RemoteCoord::close() {
  rpc_.close();
  watch_rpc_.close();
}

Result<std::string> RemoteCoord::call(uint32_t method, const std::string& req) {
  auto r = rpc_.call(method, req, 10000);
  if (!r.ok()) return ErrorCode::ETCD_ERROR;
  return r;
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

Result<int64_t> RemoteCoord::watch_prefix(const std::string& prefix, WatchCallback cb) {
  {
    std::lock_guard<std::mutex> lk(wmu_);
    if (!watch_connected_) {
      if (watch_rpc_.connect(host_, port_, 3000) != ErrorCode::OK) return ErrorCode::ETCD_WATCH_ERROR;
      watch_rpc_.enable_push([this](uint32_t topic, const std::string& payload) {
        if (topic != kTopicWatch) return;
        wire::Reader r(payload);
        const int64_t id = r.i64();
        WatchEvent ev;
        ev.type = static_cast<EventType>(r.u32());
        ev.key = r.str();
        ev.value = r.str();
        ev.revision = r.i64();
        if (!r.ok()) return;
        WatchCallback f;
        {
          std::lock_guard<std::mutex> l2(wmu_);
          auto it = watches_.find(id);
          if (it != watches_.end()) {
            f = it->second;
            ++running_[id];
            push_thread_ = std::this_thread::get_id();
          } else {
            pending_[id].push_back(ev);  // event raced ahead of the watch response
          }
        }
        if (f) {
          f(ev);
          std::lock_guard<std::mutex> l2(wmu_);
          if (--running_[id] == 0) running_.erase(id);
          wcv_.notify_all();
        }
      });
      watch_connected_ = true;
    }
  }
  wire::Writer w;
  w.str(prefix);
  auto resp = watch_rpc_.call(M_WATCH, w.data(), 10000);
  if (!resp.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  wire::Reader rd(resp.value());
  const ErrorCode ec = rd.ec();
  if (ec != ErrorCode::OK) return ec;
  const int64_t id = rd.i64();
  std::vector<WatchEvent> early;
  {
    std::lock_guard<std::mutex> lk(wmu_);
    watches_[id] = cb;
    auto it = pending_.find(id);
    if (it != pending_.end()) {
      early = std::move(it->second);
      pending_.erase(it);
    }
  }
  for (const auto& ev : early) cb(ev);
  return id;
}

ErrorCode RemoteCoord::unwatch(int64_t id) {
  {
    std::unique_lock<std::mutex> lk(wmu_);
    watches_.erase(id);
    pending_.erase(id);
    // barrier: a callback already running on the push thread finishes before we return (unless we ARE that thread)
    if (std::this_thread::get_id() != push_thread_) wcv_.wait(lk, [&] { return running_.find(id) == running_.end(); });
    if (!watch_connected_) return ErrorCode::ETCD_WATCH_ERROR;
  }
  wire::Writer w;
  w.i64(id);
  auto resp = watch_rpc_.call(M_UNWATCH, w.data(), 10000);
  if (!resp.ok()) return ErrorCode::ETCD_WATCH_ERROR;
  wire::Reader rd(resp.value());
  return rd.ec();
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
    std::lock_guard<std::mutex> lk(mu_);
    election_leases_[election] = g.value();
    is_leader = true;
  } else {
    store_->revoke_lease(g.value());
  }
  return ErrorCode::OK;
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
  if (!lease || store_->keep_alive(lease) != ErrorCode::OK) return ErrorCode::NOT_LEADER;
  auto cur = store_->get("/blackbird/elections/" + election + "/leader");
  if (!cur.ok() || cur.value() != candidate) return ErrorCode::NOT_LEADER;
  return ErrorCode::OK;
}

}  // namespace bb::coord
