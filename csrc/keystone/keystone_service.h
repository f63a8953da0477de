// KeystoneService: the control plane.
//
// Parity: reference include/blackbird/keystone/keystone_service.h:84-245 — lifecycle
// (initialize/start/stop), object API (object_exists, get_workers, put_start, put_complete,
// put_cancel, remove_object, remove_all_objects), the five batch_* calls, get_workers_info,
// get_memory_pools, remove_worker, get_cluster_stats, get_view_version; ObjectInfo (:21-57),
// WorkerInfo (:62-74); coordination schema of SURVEY §2.3; TTL GC + watermark eviction
// (keystone_service.cpp:417-451, 530-584); dead-worker cleanup (:956-1004).
//
// Re-designed (every item is a reference defect in SURVEY §2.7/§2.8):
//  * explicit PENDING -> COMPLETE state machine; get_workers never serves in-flight puts (#2)
//  * object table sharded 32 ways, no global write lock, no O(N) debug dump per put (#1, hard part 7)
//  * expired keys are reclaimed inline by put_start (#4); eviction is LRU by last access and
//    frees allocator ranges (#5); remove_all_objects frees ranges too
//  * utilisation comes from the allocator's live accounting, so the watermark actually fires (#3)
//  * tier demotion (GPU -> DRAM -> CXL -> NVMe) through a pluggable mover instead of dropping
//  * dead workers invalidate the copies they held; surviving replicas keep serving and lost
//    copies are re-replicated through the mover; objects with no survivor are removed
//  * real leader election (CAS + lease) with a standby that resumes from the object WAL
//  * client sessions (register / ping with TTL) and Prometheus metrics are real
#pragma once
#include <array>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common/durable_log.h"
#include "common/spin.h"
#include "alloc/allocator.h"
#include "common/metrics.h"
#include "common/tenant.h"
#include "common/types.h"
#include "coord/coord.h"

namespace bb::keystone {

enum class ObjectState : uint32_t { PENDING = 0, COMPLETE = 1 };
using ShardTokens = std::vector<std::vector<std::string>>;  // [copy][shard] reservation token ids ("" = none)

struct ObjectInfo {
  ObjectKey key;
  size_t size = 0;
  TimePoint created;
  TimePoint last_accessed;
  WorkerConfig config;
  std::vector<CopyPlacement> copies;
  ObjectState state = ObjectState::PENDING;
  std::string owner_client;  // session that started the put
  std::string tenant;        // the tenant it is charged to (common/tenant.h); "" = put by a member
  std::vector<std::string> extra_ledgers;  // allocator ledger keys besides `key` (repair / demotion)
  uint32_t reads_below_top = 0;            // reads served while the object sat on a lower tier (promotion policy)
  ShardTokens tokens;       // reservation tokens of a PENDING object (reservation protocol); cleared once committed
  bool reserved = false;    // its shards are reserved / committed at the workers
  bool committing = false;  // put_complete has logged COMPLETE but not yet flipped `state` (snapshots count it as COMPLETE)

  uint64_t ttl_ms() const { return config.ttl_ms; }
  bool is_expired(TimePoint now = Clock::now()) const {
    return config.ttl_ms != 0 && now - created > std::chrono::milliseconds(config.ttl_ms);
  }
  bool has_complete_copy() const { return state == ObjectState::COMPLETE && !copies.empty(); }
  void touch() { last_accessed = Clock::now(); }
};

struct WorkerInfo {
  WorkerId worker_id;
  NodeId node_id;
  std::string endpoint;
  TimePoint last_heartbeat;
  std::vector<MemoryPoolId> pools;
  WorkerRecord record;
  bool is_stale(std::chrono::seconds ttl, TimePoint now = Clock::now()) const { return now - last_heartbeat > ttl; }
};

struct PutStartItem {
  ObjectKey key;
  size_t size = 0;
  WorkerConfig config;
};
// Per-object shard digests reported by the writer at put_complete: [copy][shard].
using ShardChecksums = std::vector<std::vector<uint64_t>>;

// Reservation protocol (reference storage_backend.h:46-126: reserve_shard -> commit_shard | abort_shard, free_shard; SURVEY
// section 2.7 "must be real").  With `enable_reservations` the Keystone, which decides the placement, has every shard
// reserved at its worker before the writer sees it (token with an expiry), commits the tokens when the writer reports
// completion, aborts them on cancel, and frees committed shards on removal; a writer that vanishes is cleaned up by the
// WORKER (token expiry -> coordination key -> the Keystone drops the PENDING object), not by a Keystone GC pass.
struct ReservationHooks {
  std::function<ErrorCode(const ObjectKey& key, const std::vector<CopyPlacement>& copies, uint64_t ttl_ms, ShardTokens* tokens)> reserve;
  std::function<ErrorCode(const std::vector<CopyPlacement>& copies, const ShardTokens& tokens)> commit, abort;
  std::function<ErrorCode(const std::vector<CopyPlacement>& copies)> release;
  explicit operator bool() const { return reserve && commit && abort && release; }
};

// Moves the bytes of one copy to another placement (tier demotion / re-replication).  Provided
// by the data plane (worker or client library); returns OK when dst holds a verified copy and
// fills dst shard checksums.
using CopyMover = std::function<ErrorCode(const ObjectKey& key, const CopyPlacement& src, CopyPlacement& dst, ChecksumAlgo algo)>;

// Scrub: re-hashes one stored copy where it lies and compares with the recorded digests (client/copy_mover.h).
using CopyVerifier = std::function<ErrorCode(const ObjectKey& key, const CopyPlacement& copy, ChecksumAlgo algo)>;

// What a tenant (common/tenant.h) holds right now and what it may hold.
struct TenantUsage {
  std::string name;
  uint64_t used_bytes = 0;  // size x replication_factor of every object it put and that still exists
  uint64_t objects = 0;
  uint64_t quota_bytes = 0;  // 0 = unlimited (from the tenant table; 0 as well for a tenant that left the table)
  uint64_t max_objects = 0;
};

struct ScrubReport {
  uint64_t objects = 0;        // COMPLETE objects looked at
  uint64_t copies = 0;         // copies hashed
  uint64_t corrupt = 0;        // copies whose bytes no longer match their digest
  uint64_t healed = 0;         // ... replaced by a fresh copy made from a healthy replica
  uint64_t unrecoverable = 0;  // objects with no healthy copy left (still listed; reads fail their digest check)
  uint64_t unreachable = 0;    // copies whose worker could not be asked (left to failure detection)
};

class KeystoneService {
 public:
  explicit KeystoneService(const KeystoneConfig& config, std::shared_ptr<coord::CoordService> coord = nullptr);
  ~KeystoneService();
  KeystoneService(const KeystoneService&) = delete;
  KeystoneService& operator=(const KeystoneService&) = delete;

  ErrorCode initialize();
  ErrorCode start();
  void stop();
  bool is_running() const noexcept { return running_.load(); }
  // In HA mode leadership is also bounded in *local* time: it lapses on its own `margin` before the election lease
  // can expire at the store, so a leader that was paused (SIGSTOP, VM freeze, long GC) comes back as a non-leader and
  // cannot acknowledge anything until it has talked to the store again.  WAL writes are additionally fenced by term.
  bool is_leader() const noexcept {
    if (!config_.enable_ha) return true;
    return leader_.load(std::memory_order_acquire) &&
           Clock::now().time_since_epoch().count() < lease_deadline_.load(std::memory_order_acquire);
  }
  int64_t leader_term() const noexcept { return term_.load(); }
  const KeystoneConfig& config() const { return config_; }

  // ---- object API
  Result<bool> object_exists(const ObjectKey& key);
  Result<std::vector<CopyPlacement>> get_workers(const ObjectKey& key);
  Result<std::vector<CopyPlacement>> put_start(const ObjectKey& key, size_t data_size, const WorkerConfig& config,
                                               const std::string& client_id = "", const std::string& client_node = "");
  ErrorCode put_complete(const ObjectKey& key);
  ErrorCode put_complete(const ObjectKey& key, const ShardChecksums& checksums);
  ErrorCode put_cancel(const ObjectKey& key);
  ErrorCode remove_object(const ObjectKey& key);
  Result<size_t> remove_all_objects();

  // ---- batch API (one lock acquisition per shard group, per-item results)
  std::vector<Result<bool>> batch_object_exists(const std::vector<ObjectKey>& keys);
  std::vector<Result<std::vector<CopyPlacement>>> batch_get_workers(const std::vector<ObjectKey>& keys);
  std::vector<Result<std::vector<CopyPlacement>>> batch_put_start(const std::vector<PutStartItem>& items,
                                                                  const std::string& client_id = "",
                                                                  const std::string& client_node = "");
  std::vector<ErrorCode> batch_put_complete(const std::vector<ObjectKey>& keys);
  std::vector<ErrorCode> batch_put_complete(const std::vector<ObjectKey>& keys, const std::vector<ShardChecksums>& checksums);
  std::vector<ErrorCode> batch_put_cancel(const std::vector<ObjectKey>& keys);
  std::vector<ErrorCode> batch_remove_object(const std::vector<ObjectKey>& keys);

  // ---- cluster / admin
  Result<ClusterStats> get_cluster_stats() const;
  ViewVersionId get_view_version() const noexcept { return view_version_.load(); }
  ErrorCode get_workers_info(std::vector<WorkerInfo>& out) const;
  ErrorCode get_memory_pools(std::vector<MemoryPool>& out) const;
  ErrorCode remove_worker(const WorkerId& id);
  Result<ObjectInfo> get_object_info(const ObjectKey& key) const;
  // Keys of COMPLETE, unexpired objects starting with `prefix`, in lexicographic order, at most `limit` (0 = 10000)
  // strictly after `start_after` (pagination).  An extension: the reference has no listing call.
  struct ListedObject {
    ObjectKey key;
    uint64_t size = 0;
    uint32_t copies = 0;
    StorageClass tier = StorageClass::STORAGE_UNSPECIFIED;  // tier of the first shard of copy 0
  };
  // Pool compaction (a roadmap item of the reference, README.md:146-153): re-places the objects that sit highest in
  // `pool` into lower holes of the same tier (best fit), through the installed mover (digest re-checked, placements
  // swapped atomically, old extents freed), until `max_moves` objects moved or nothing moves any more.  Returns the
  // number of objects moved.  NOT_IMPLEMENTED without a mover, MEMORY_POOL_NOT_FOUND for an unknown pool.
  Result<size_t> compact_pool(const MemoryPoolId& pool, size_t max_moves = 64);
  // Decommission without losing redundancy: the worker's pools stop receiving placements, every object with a shard there
  // is re-placed on the other workers (same tier first, bytes through the mover, digests re-checked) while its old copies
  // keep serving, and only then is the worker removed.  Returns the number of objects moved; objects that found no room
  // elsewhere stay where they are, the worker stays registered (still draining) and the call reports INSUFFICIENT_SPACE.
  Result<size_t> drain_worker(const WorkerId& id);
  // One round of the automatic trigger (health loop): every pool above `compaction_fragmentation_threshold` gets up to
  // 8 moves.  Returns the number of objects moved.
  size_t run_compaction_once();
  std::vector<ListedObject> list_objects(const std::string& prefix, size_t limit = 0, const std::string& start_after = "") const;

  // Direct registration (in-process deployments and tests; the coordination watchers call
  // the same functions).
  ErrorCode register_worker(const WorkerRecord& rec);
  ErrorCode register_memory_pool(const MemoryPool& pool);
  ErrorCode worker_heartbeat(const WorkerId& id);
  void handle_worker_death(const WorkerId& id);
  // A single pool left the registry (deregistration event): same invalidation as a death, limited to that pool.
  void handle_pool_removed(const MemoryPoolId& pool);

  // ---- client sessions (reference README "client heartbeats"; absent in its code)
  Result<std::string> client_register(const std::string& node_id);
  Result<ViewVersionId> client_ping(const std::string& client_id);

  // ---- tiering / repair
  void set_copy_mover(CopyMover m);
  // Installs the transport of the reservation protocol (client/copy_mover.h: the workers' data servers).  Takes effect
  // when `enable_reservations` is set in the configuration.
  void set_reservation_hooks(ReservationHooks h);
  bool reservations_enabled() const { return config_.enable_reservations && reservations_on_.load(std::memory_order_relaxed); }
  // Runs one TTL sweep / eviction pass / repair pass synchronously (also used by the threads).
  size_t run_gc_once();
  size_t run_eviction_once();
  // Read-driven promotion: objects read `promote_after_reads` times on a lower tier move back to the fastest tier that has
  // room below the watermark.  Returns the number of objects promoted.
  size_t run_promotion_once();
  uint64_t tier_capacity(StorageClass sc) const;
  // Explicit tier move (promotion or demotion) of every copy of `key` to `target`; no-op when already there.
  ErrorCode migrate_object(const ObjectKey& key, StorageClass target);
  size_t run_repair_once();
  // Scrub (the reference never re-reads what it stored): every COMPLETE object under `prefix` (at most `max_objects`, 0 = all)
  // has each copy hashed by the worker that holds it.  A copy that no longer matches is replaced from a healthy replica --
  // new extents, verified bytes, the bad extents released -- so the object is never served from, or repaired out of, bit rot.
  void set_copy_verifier(CopyVerifier v);
  Result<ScrubReport> scrub(const std::string& prefix = "", size_t max_objects = 0);
  // Per-tenant holdings against their budgets (every tenant in the table, plus any that still owns objects).
  std::vector<TenantUsage> tenant_usage() const;
  void count_acl_denial() { metrics_.inc("tenant_acl_denials_total"); }  // the RPC layer refused a key outside a tenant's grants
  double tier_utilization(StorageClass sc) const;

  // ---- observability
  std::string metrics_text() const;  // Prometheus exposition
  Json stats_json() const;
  alloc::AllocatorStats allocator_stats() const { return allocator_->get_allocator_stats(); }

 private:
  static constexpr size_t kShards = 128;
  // One queued metadata-log operation.  Mutations enqueue under Shard::mu (which fixes their order) and the log I/O
  // happens after that lock is dropped (flush_wal), never under the spin lock.
  struct WalOp {
    ObjectKey key;
    std::string value;          // encoded ObjectInfo; empty + del = tombstone
    bool del = false;
    ErrorCode* result = nullptr;  // optional: where the enqueuing thread wants the write's outcome (it outlives the flush)
  };
  struct HookOp {
    std::vector<CopyPlacement> copies;
    ShardTokens tokens;
    bool release = false;  // committed shards of a COMPLETE object (free) vs. tokens of a PENDING one (abort)
  };
  struct Shard {
    mutable SpinMutex mu;  // sub-microsecond sections taken by every client on every object: never sleep on it
    std::unordered_map<ObjectKey, ObjectInfo> objects;
    std::vector<WalOp> wal_queue;  // guarded by mu
    std::vector<HookOp> hook_queue;  // guarded by mu: reservations to abort / release once the lock is dropped
    std::mutex wal_mu;             // serialises the writers of this shard's log records (order = queue order)
  };
  // unique_lock on a shard that flushes the shard's queued log records when it goes out of scope (after unlocking).
  class ShardGuard {
   public:
    ShardGuard(KeystoneService* ks, Shard& sh) : ks_(ks), sh_(sh), lk_(sh.mu) {}
    ~ShardGuard() { finish(); }
    void finish() {
      if (!lk_.owns_lock()) return;
      const bool dirty = !sh_.wal_queue.empty();
      std::vector<HookOp> hooks;
      hooks.swap(sh_.hook_queue);
      lk_.unlock();
      if (dirty) ks_->flush_wal(sh_);
      if (!hooks.empty()) ks_->run_hook_ops(hooks);
    }
    void relock() { lk_.lock(); }
   private:
    KeystoneService* ks_;
    Shard& sh_;
    std::unique_lock<SpinMutex> lk_;
  };
  friend class ShardGuard;
  Shard& shard_for(const ObjectKey& key) { return shards_[std::hash<ObjectKey>{}(key) % kShards]; }
  const Shard& shard_for(const ObjectKey& key) const { return shards_[std::hash<ObjectKey>{}(key) % kShards]; }

  ErrorCode setup_coordination();
  void load_existing_state();
  void on_worker_event(const std::string& key, const std::string& value, bool is_delete);
  void on_heartbeat_event(const std::string& key, const std::string& value, bool is_delete);
  void on_legacy_pool_event(const std::string& key, const std::string& value, bool is_delete);
  void gc_loop();
  void health_loop();
  void keepalive_loop();
  // Metadata log.  Sinks: the coordination store (HA: shared with the standby, every write fenced by the leader's
  // term) or a local DurableLog under `wal_path` (single Keystone: survives its own restart).  persist / unpersist are
  // called with the shard lock held and only enqueue.
  bool wal_enabled() const { return wal_on_.load(std::memory_order_relaxed); }
  void persist_object(Shard& sh, const ObjectInfo& info, ErrorCode* result = nullptr);
  void unpersist_object(Shard& sh, const ObjectKey& key);
  void flush_wal(Shard& sh);
  ErrorCode wal_write(const WalOp& op);
  void recover_objects_from_wal();
  void recover_objects(std::vector<std::pair<std::string, std::string>>& records);
  ErrorCode open_local_wal();
  void maybe_snapshot_local_wal();
  void reset_object_state();  // drops the object table and the allocator (leadership lost / about to be rebuilt)
  void step_down(const char* why);
  void run_hook_ops(const std::vector<HookOp>& ops);
  ErrorCode reserve_after_start(const ObjectKey& key, Result<std::vector<CopyPlacement>>& placed);
  void on_reservation_expired(const std::string& key, const std::string& owner, bool is_delete);
  void arm_lease_deadline(TimePoint refreshed_at);
  std::string election_name() const { return "keystone-" + config_.cluster_id; }
  void bump_view() { view_version_.fetch_add(1); }
  ErrorCode erase_locked(Shard& sh, const ObjectKey& key, bool free_ranges);
  std::vector<CopyPlacement> live_copies(const ObjectInfo& info) const;
  bool pool_alive(const MemoryPoolId& id) const;
  void invalidate_pools(const std::vector<MemoryPoolId>& dead, const std::string& why);
  std::string cluster_prefix() const { return "/blackbird/clusters/" + config_.cluster_id + "/"; }
  bool interruptible_sleep(std::chrono::milliseconds d);

  KeystoneConfig config_;
  std::shared_ptr<coord::CoordService> coord_;
  std::unique_ptr<alloc::KeystoneAllocatorAdapter> allocator_;
  std::array<Shard, kShards> shards_;

  mutable std::shared_mutex pools_mu_;  // lock order: workers_mu_ -> pools_mu_ -> shard.mu
  std::unordered_map<MemoryPoolId, MemoryPool> pools_;
  mutable std::shared_mutex workers_mu_;
  std::unordered_map<WorkerId, WorkerInfo> workers_;

  mutable std::mutex clients_mu_;
  struct ClientSession {
    std::string node_id;
    TimePoint last_ping;
  };
  std::unordered_map<std::string, ClientSession> clients_;

  std::mutex mover_mu_;
  CopyMover mover_;
  CopyVerifier verifier_;
  std::vector<MemoryPoolId> draining_;  // guarded by pools_mu_: pools of workers being drained (no new placements)
  ReservationHooks res_hooks_;  // guarded by mover_mu_ (copied out before use)
  std::atomic<bool> reservations_on_{false};
  void refresh_top_tier_locked();  // caller holds pools_mu_ exclusively
  std::atomic<int> top_tier_rank_{-1};  // rank of the fastest tier present (promotion policy fast check)
  std::mutex promo_mu_;
  std::vector<ObjectKey> promo_queue_;
  // items[first, last) share one size and one WorkerConfig: place them as a run (one ranking, one allocator call per chunk).
  // False = the policy is not a one-shard-per-object one; nothing was done.  Caller holds pools_mu_ (shared).
  bool put_start_run(const std::vector<PutStartItem>& items, size_t first, size_t last, const std::string& client_id,
                     const std::string& client_node, std::vector<Result<std::vector<CopyPlacement>>>& out);
  Result<std::vector<CopyPlacement>> put_start_locked(const ObjectKey& key, size_t data_size, const WorkerConfig& config,
                                                      const std::string& client_id, const std::string& client_node);
  struct HotMetrics {
    std::atomic<uint64_t>*put_start_total = nullptr, *put_start_failed_total = nullptr, *put_bytes_total = nullptr,
                         *put_complete_total = nullptr, *get_workers_total = nullptr, *remove_total = nullptr;
    Histogram* put_start_latency = nullptr;
  } hot_;
  // `accept` (optional) sees the old and the freshly allocated placements before any byte moves; false = roll back.
  using PlacementFilter = std::function<bool(const std::vector<CopyPlacement>& old_copies, const std::vector<CopyPlacement>& fresh)>;
  Result<ScrubReport> scrub_from(const std::string& prefix, size_t max_objects, size_t& cursor);
  Result<CopyPlacement> place_extra_copy(const ObjectInfo& o, size_t skip, std::string& ledger);
  ErrorCode replace_copy(const CopyMover& mover, const ObjectInfo& o, size_t idx, const CopyPlacement& source);
  size_t scrub_cursor_ = 0;  // health loop only
  // Admission control per tenant: charged in put_start, released wherever an object leaves the table (erase_locked).
  struct TenantCount {
    uint64_t bytes = 0, objects = 0;
  };
  mutable std::mutex tenant_mu_;
  std::unordered_map<std::string, TenantCount> tenant_usage_;
  bool tenant_admit(const Tenant& t, uint64_t bytes);  // false = over its budget (nothing charged)
  void tenant_charge(const std::string& name, uint64_t bytes);  // recovery: no check
  void tenant_release(const std::string& name, uint64_t bytes);
  static uint64_t tenant_charge_of(const ObjectInfo& o) { return static_cast<uint64_t>(o.size) * std::max<size_t>(1, o.config.replication_factor); }
  ErrorCode migrate_with(const CopyMover& mover, const ObjectKey& key, const std::vector<StorageClass>& targets,
                         const PlacementFilter& accept = nullptr);

  std::atomic<bool> running_{false};
  std::atomic<bool> leader_{false};
  std::atomic<int64_t> term_{0};                 // fencing token of the current leadership (0 = none)
  std::atomic<int64_t> lease_deadline_{0};       // Clock ticks; is_leader() is false from here on until re-armed
  std::atomic<bool> wal_on_{false};
  std::unique_ptr<DurableLog> local_wal_;
  std::mutex local_snap_mu_;
  // recovered objects whose pools had not registered yet: pool id -> keys (their extents are adopted on registration)
  std::mutex unadopted_mu_;
  std::unordered_map<MemoryPoolId, std::vector<ObjectKey>> unadopted_;
  std::atomic<ViewVersionId> view_version_{0};
  std::mutex sleep_mu_;
  std::condition_variable sleep_cv_;
  std::thread gc_thread_, health_thread_, keepalive_thread_;
  std::vector<int64_t> watch_ids_;  // coordination watches owned by this service (released in stop(), which is a barrier)
  std::string candidate_id_;

  mutable Metrics metrics_;
};

}  // namespace bb::keystone
