// Coordination substrate: the role etcd plays for the reference.
//
// Parity: reference include/blackbird/etcd/etcd_service.h:30-245 — KV get/put/del, prefix list,
// TTL puts, leases (grant / keep_alive / revoke / refresh), prefix + key watches, service
// registry helpers and leader election (campaign / get_leader / resign; `campaign_leader` is a
// stub in the reference, etcd_service.cpp:379-385 — it is real here).
//
// etcd itself is unavailable offline, so the store is implemented in-tree behind `CoordStore`:
//   MemCoord    — in-process, revisioned KV + leases + watches (+ txn put-if-absent / CAS).
//   CoordServer — serves a MemCoord over the framed RPC protocol (the `bb-coord` daemon).
//   RemoteCoord — client of a CoordServer; watch events arrive as server push frames.
// The etcd key schema of SURVEY §2.3 is preserved so a real etcd adapter could drop in.
// Fixes over the reference: TTL puts reuse one lease per key (no lease leak, bug #12), watch
// callbacks never run under the store lock, all endpoints are tried on connect (bug: first only).
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common/durable_log.h"
#include "common/types.h"
#include "net/tcp.h"

namespace bb::coord {

struct KeyValue {
  std::string key;
  std::string value;
  int64_t create_revision = 0;
  int64_t mod_revision = 0;
  LeaseId lease = 0;
};

enum class EventType : uint32_t { PUT = 0, DELETE = 1 };
struct WatchEvent {
  EventType type = EventType::PUT;
  std::string key;
  std::string value;  // new value for PUT, last value for DELETE
  int64_t revision = 0;
};
using WatchCallback = std::function<void(const WatchEvent&)>;

class CoordStore {
 public:
  virtual ~CoordStore() = default;
  virtual ErrorCode put(const std::string& key, const std::string& value, LeaseId lease = 0) = 0;
  virtual Result<KeyValue> get_kv(const std::string& key) = 0;  // ETCD_KEY_NOT_FOUND
  virtual ErrorCode del(const std::string& key) = 0;            // OK even when absent
  virtual Result<std::vector<KeyValue>> get_with_prefix(const std::string& prefix) = 0;
  virtual Result<size_t> del_prefix(const std::string& prefix) = 0;
  virtual Result<LeaseId> grant_lease(int64_t ttl_sec) = 0;
  virtual ErrorCode keep_alive(LeaseId lease) = 0;  // ETCD_LEASE_ERROR when expired/unknown
  virtual ErrorCode revoke_lease(LeaseId lease) = 0;
  virtual Result<int64_t> lease_remaining_ms(LeaseId lease) = 0;
  // txn: create only if the key does not exist.  value() == true when this call created it.
  virtual Result<bool> put_if_absent(const std::string& key, const std::string& value, LeaseId lease = 0) = 0;
  // txn: replace / delete only if the current value equals `expected`.
  virtual Result<bool> compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                        LeaseId lease = 0) = 0;
  virtual Result<bool> compare_and_delete(const std::string& key, const std::string& expected) = 0;
  // Fenced writes (etcd: txn{compare: create_revision(guard_key) == guard_create_revision; success: put | delete}).
  // The guard is the election key and its create revision is the leader's term: a deposed leader's writes fail
  // (value() == false) no matter how late they arrive.
  virtual Result<bool> guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                                   const std::string& value) = 0;
  virtual Result<bool> guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) = 0;
  virtual Result<int64_t> watch_prefix(const std::string& prefix, WatchCallback cb) = 0;
  virtual ErrorCode unwatch(int64_t watch_id) = 0;
  virtual int64_t revision() = 0;

  Result<std::string> get(const std::string& key) {
    auto kv = get_kv(key);
    if (!kv.ok()) return kv.error();
    return kv.value().value;
  }
};

// ---------------------------------------------------------------- in-process store
class MemCoord : public CoordStore {
 public:
  MemCoord();
  ~MemCoord() override;
  ErrorCode put(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<KeyValue> get_kv(const std::string& key) override;
  ErrorCode del(const std::string& key) override;
  Result<std::vector<KeyValue>> get_with_prefix(const std::string& prefix) override;
  Result<size_t> del_prefix(const std::string& prefix) override;
  Result<LeaseId> grant_lease(int64_t ttl_sec) override;
  ErrorCode keep_alive(LeaseId lease) override;
  ErrorCode revoke_lease(LeaseId lease) override;
  Result<int64_t> lease_remaining_ms(LeaseId lease) override;
  Result<bool> put_if_absent(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<bool> compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                LeaseId lease = 0) override;
  Result<bool> compare_and_delete(const std::string& key, const std::string& expected) override;
  Result<bool> guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                           const std::string& value) override;
  Result<bool> guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) override;
  Result<int64_t> watch_prefix(const std::string& prefix, WatchCallback cb) override;
  ErrorCode unwatch(int64_t watch_id) override;
  int64_t revision() override;

  // Durability (`bb-coord --data-dir`): every mutation is appended to a log in `dir` and fdatasync'ed (group commit)
  // before the call returns; the log is compacted into snapshots.  Re-opening the directory restores keys, revisions
  // and leases (re-armed with their full TTL, like etcd after a restart; lease ids are preserved so that clients keep
  // refreshing the leases they hold).  Call before the store is shared.
  ErrorCode open_durable(const std::string& dir, bool fsync = true, uint64_t snapshot_bytes = 64ull << 20);
  bool durable() const { return log_ != nullptr; }
  uint64_t recovered_records() const { return recovered_records_; }

  // Test / fault-injection hooks: moves the store's clock forward (expires leases) and blocks
  // until the resulting events have been delivered.
  void advance_time_ms(int64_t ms);
  void flush_events();
  size_t lease_count();
  size_t key_count();

 private:
  struct Lease {
    int64_t ttl_ms;
    int64_t expires_at_ms;
    std::set<std::string> keys;
  };
  struct Watcher {
    std::string prefix;
    WatchCallback cb;
    int running = 0;  // invocations in flight (guarded by wmu_); unwatch() waits for 0
  };
  std::condition_variable wcv_;  // signalled when a watcher's invocation returns
  int64_t now_ms() const;
  // durable log plumbing: *_locked mutators append under mu_ and leave the record's sequence number in pending_seq_;
  // the public entry point syncs it after dropping mu_ (commit()).
  void log_locked(const std::string& rec);
  uint64_t take_seq_locked() {
    const uint64_t s = pending_seq_;
    pending_seq_ = 0;
    return s;
  }
  ErrorCode commit(uint64_t seq, ErrorCode ec);
  std::string snapshot_locked() const;
  void load_snapshot(const std::string& blob);
  void apply_record(std::string_view rec);
  std::unique_ptr<DurableLog> log_;
  uint64_t pending_seq_ = 0;
  bool replaying_ = false;
  uint64_t recovered_records_ = 0;
  std::mutex snap_mu_;  // one snapshot at a time
  ErrorCode put_locked(const std::string& key, const std::string& value, LeaseId lease);
  bool del_locked(const std::string& key);
  void emit_locked(EventType t, const std::string& key, const std::string& value);
  void expire_locked();
  void expiry_loop();
  void dispatch_loop();

  std::mutex mu_;
  std::map<std::string, KeyValue> kv_;
  std::unordered_map<LeaseId, Lease> leases_;
  int64_t revision_ = 1;
  LeaseId next_lease_ = 1000;
  std::atomic<int64_t> clock_offset_ms_{0};
  // watchers + event queue (delivered by dispatch thread, never under mu_)
  std::mutex wmu_;
  std::map<int64_t, std::shared_ptr<Watcher>> watchers_;
  int64_t next_watch_ = 1;
  std::mutex qmu_;
  std::condition_variable qcv_;
  std::condition_variable qidle_;
  std::deque<WatchEvent> queue_;
  bool dispatching_ = false;
  std::atomic<bool> stop_{false};
  std::condition_variable expiry_cv_;
  std::thread expiry_thread_;
  std::thread dispatch_thread_;
};

// ---------------------------------------------------------------- daemon + remote client
class CoordServer {
 public:
  explicit CoordServer(std::shared_ptr<MemCoord> store = nullptr);
  ~CoordServer();
  ErrorCode start(const std::string& host, uint16_t port);
  void stop();
  uint16_t port() const { return rpc_.port(); }
  std::shared_ptr<MemCoord> store() { return store_; }

 private:
  struct ConnState;
  std::shared_ptr<MemCoord> store_;
  net::RpcServer rpc_;
};

class RemoteCoord : public CoordStore {
 public:
  RemoteCoord() = default;
  ~RemoteCoord() override;
  // endpoints: comma separated "host:port" list; every endpoint is tried in order.
  ErrorCode connect(const std::string& endpoints, int timeout_ms = 3000);
  void close();
  ErrorCode put(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<KeyValue> get_kv(const std::string& key) override;
  ErrorCode del(const std::string& key) override;
  Result<std::vector<KeyValue>> get_with_prefix(const std::string& prefix) override;
  Result<size_t> del_prefix(const std::string& prefix) override;
  Result<LeaseId> grant_lease(int64_t ttl_sec) override;
  ErrorCode keep_alive(LeaseId lease) override;
  ErrorCode revoke_lease(LeaseId lease) override;
  Result<int64_t> lease_remaining_ms(LeaseId lease) override;
  Result<bool> put_if_absent(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<bool> compare_and_swap(const std::string& key, const std::string& expected, const std::string& value,
                                LeaseId lease = 0) override;
  Result<bool> compare_and_delete(const std::string& key, const std::string& expected) override;
  Result<bool> guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                           const std::string& value) override;
  Result<bool> guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) override;
  Result<int64_t> watch_prefix(const std::string& prefix, WatchCallback cb) override;
  ErrorCode unwatch(int64_t watch_id) override;
  int64_t revision() override;
  uint64_t reconnects() const { return reconnects_.load(); }

 private:
  struct WatchReg {
    std::string prefix;
    WatchCallback cb;
    int64_t server_id = 0;         // id of the current incarnation of this watch on the server
    std::set<std::string> known;   // keys this watcher has been told exist (to synthesise DELETEs after a reconnect)
  };
  Result<std::string> call(uint32_t method, const std::string& req);
  bool connect_any(net::RpcClient& c, int timeout_ms);
  bool open_watch_channel();
  Result<int64_t> server_watch(const std::string& prefix);
  void deliver(int64_t local_id, const WatchEvent& ev);
  void monitor_loop();
  net::RpcClient rpc_;        // request/response
  net::RpcClient watch_rpc_;  // push channel
  std::mutex conn_mu_;        // reconnect of rpc_
  std::mutex watch_conn_mu_;  // watch channel (re)establishment, M_WATCH / M_UNWATCH calls
  std::mutex wmu_;
  std::map<int64_t, WatchReg> watches_;   // local id -> registration
  std::map<int64_t, int64_t> by_server_;  // server id -> local id
  int64_t next_local_watch_ = 1;
  std::map<int64_t, int> running_;  // callback invocations in flight per watch (guarded by wmu_)
  std::condition_variable wcv_;
  std::map<int64_t, std::vector<WatchEvent>> pending_;  // by server id
  bool watch_connected_ = false;
  std::vector<std::pair<std::string, uint16_t>> endpoints_;
  std::thread monitor_;
  std::atomic<bool> closing_{false};
  std::atomic<uint64_t> reconnects_{0};
  std::atomic<uint64_t> conn_gen_{0};
};

// ---------------------------------------------------------------- service facade (EtcdService parity)
class CoordService {
 public:
  // endpoints == "" or "mem://" -> private in-process store; "mem://<name>" -> process-wide
  // shared in-process store (in-proc keystone + workers); otherwise RemoteCoord.
  explicit CoordService(const std::string& endpoints);
  explicit CoordService(std::shared_ptr<CoordStore> store);
  ~CoordService();

  ErrorCode connect();
  bool is_connected() const { return connected_; }
  std::shared_ptr<CoordStore> store() { return store_; }

  ErrorCode get(const std::string& key, std::string& value);
  ErrorCode put(const std::string& key, const std::string& value);
  // One lease per key is reused and refreshed (reference leaks a lease per call).
  ErrorCode put_with_ttl(const std::string& key, const std::string& value, int64_t ttl_sec);
  ErrorCode del(const std::string& key);
  ErrorCode get_with_prefix(const std::string& prefix, std::vector<std::string>& keys, std::vector<std::string>& values);
  ErrorCode grant_lease(int64_t ttl_sec, LeaseId& lease);
  ErrorCode put_with_lease(const std::string& key, const std::string& value, LeaseId lease);
  ErrorCode keep_alive(LeaseId lease);
  ErrorCode revoke_lease(LeaseId lease);
  ErrorCode refresh_lease(LeaseId lease) { return keep_alive(lease); }
  using WatchCb = std::function<void(const std::string& key, const std::string& value, bool is_delete)>;
  // `watch_id` (optional) receives the id to pass to unwatch().  unwatch() is a barrier: when it returns, the callback
  // is not running and will not run again, so the owner of the captured state may be destroyed.
  ErrorCode watch_prefix(const std::string& prefix, WatchCb cb, int64_t* watch_id = nullptr);
  ErrorCode unwatch(int64_t watch_id);
  ErrorCode watch_key(const std::string& key, WatchCb cb);
  ErrorCode unwatch_key(const std::string& key);
  // service registry: /blackbird/services/<name>/<id> -> address
  ErrorCode register_service(const std::string& name, const std::string& id, const std::string& address, int64_t ttl_sec);
  ErrorCode discover_service(const std::string& name, std::vector<std::string>& addresses);
  ErrorCode unregister_service(const std::string& name, const std::string& id);
  // leader election: /blackbird/elections/<name>/leader (CAS create with a lease)
  ErrorCode campaign_leader(const std::string& election, const std::string& candidate, int64_t ttl_sec, bool& is_leader);
  ErrorCode get_leader(const std::string& election, std::string& leader);
  ErrorCode resign_leader(const std::string& election, const std::string& candidate);
  // Keeps this candidate's leadership lease alive; NOT_LEADER when it was lost.
  ErrorCode refresh_leadership(const std::string& election, const std::string& candidate);
  // Fencing.  The term of a leadership is the create revision of the election key (monotonic across elections, as in
  // etcd's election recipe); 0 = this service does not hold the election.  fenced_put / fenced_del apply only while
  // that very incarnation of the key exists: NOT_LEADER when the guard fails (a deposed leader cannot write).
  int64_t leader_term(const std::string& election);
  ErrorCode fenced_put(const std::string& election, const std::string& key, const std::string& value);
  ErrorCode fenced_del(const std::string& election, const std::string& key);

 private:
  std::string endpoints_;
  std::shared_ptr<CoordStore> store_;
  bool connected_ = false;
  std::mutex mu_;
  std::unordered_map<std::string, LeaseId> ttl_leases_;       // key -> lease (put_with_ttl)
  std::unordered_map<std::string, LeaseId> election_leases_;  // election -> lease
  std::unordered_map<std::string, int64_t> election_terms_;   // election -> create revision of the key we own
  std::unordered_map<std::string, int64_t> key_watches_;
  std::vector<int64_t> watch_ids_;
};

// Process-wide named in-memory stores ("mem://name").
std::shared_ptr<MemCoord> shared_mem_coord(const std::string& name);
void drop_shared_mem_coord(const std::string& name);

}  // namespace bb::coord
