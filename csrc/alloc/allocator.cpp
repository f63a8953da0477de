#include "alloc/allocator.h"

#include <algorithm>
#include <unordered_set>

#include "common/log.h"

namespace bb::alloc {

// ================================================================ PoolAllocator
namespace {
uint32_t rkey_from_hex(const std::string& hex) {
  auto bytes = hex_to_bytes(hex);
  if (!bytes || bytes->empty()) return 0;
  if (bytes->size() <= 4) {
    uint32_t v = 0;
    for (uint8_t b : *bytes) v = (v << 8) | b;
    return v;
  }
  return crc32c(bytes->data(), bytes->size());  // long fabric handles: stable 32-bit tag
}
}  // namespace

PoolAllocator::PoolAllocator(const MemoryPool& pool, uint64_t align)
    : pool_id_(pool.id),
      storage_class_(pool.storage_class),
      node_id_(pool.node_id),
      base_addr_(pool.ucx_remote_addr ? pool.ucx_remote_addr : pool.base_addr),
      rkey_(rkey_from_hex(pool.ucx_rkey_hex)),
      pool_size_(pool.size),
      align_(align ? align : 1) {
  const uint64_t usable = pool.size / align_ * align_;
  if (usable) insert_free(0, usable);
  if (!pool.ucx_endpoint.empty()) {
    auto hp = split_host_port(pool.ucx_endpoint);
    if (hp) {
      endpoint_.ip = hp->first;
      endpoint_.port = hp->second;
    } else {
      reg_valid_ = false;
    }
  }
  if (!pool.ucx_rkey_hex.empty()) {
    auto key = hex_to_bytes(pool.ucx_rkey_hex);
    if (key) endpoint_.worker_key = std::move(*key);
    else reg_valid_ = false;
  }
}

void PoolAllocator::insert_free(uint64_t off, uint64_t len) {
  by_offset_[off] = len;
  by_size_.insert({len, off});
  free_bytes_ += len;
  publish();
}

void PoolAllocator::erase_free(std::map<uint64_t, uint64_t>::iterator it) {
  by_size_.erase({it->second, it->first});
  free_bytes_ -= it->second;
  by_offset_.erase(it);
  publish();
}

// Placement ranks every pool on every put: the two numbers it needs are mirrored in atomics so ranking never
// touches a pool's mutex (8 clients ranking 8 pools used to serialise on the busiest pool's lock).
void PoolAllocator::publish() {
  free_pub_.store(free_bytes_, std::memory_order_relaxed);
  largest_pub_.store(by_size_.empty() ? 0 : by_size_.rbegin()->first, std::memory_order_relaxed);
}

std::optional<Range> PoolAllocator::allocate(uint64_t size, bool prefer_best_fit) {
  const uint64_t need = aligned(size);
  std::lock_guard<SpinMutex> lk(mu_);
  if (need == 0) return Range(0, 0);
  std::map<uint64_t, uint64_t>::iterator it = by_offset_.end();
  if (prefer_best_fit) {
    auto s = by_size_.lower_bound({need, 0});  // tightest hole; lowest offset among equals
    if (s == by_size_.end()) return std::nullopt;
    it = by_offset_.find(s->second);
  } else {
    for (auto f = by_offset_.begin(); f != by_offset_.end(); ++f)
      if (f->second >= need) {
        it = f;
        break;
      }
    if (it == by_offset_.end()) return std::nullopt;
  }
  const uint64_t off = it->first, len = it->second;
  erase_free(it);
  if (len > need) insert_free(off + need, len - need);
  return Range(off, need);
}

bool PoolAllocator::allocate_at(uint64_t offset, uint64_t size) {
  const uint64_t need = aligned(size);
  if (need == 0) return true;
  if (offset % align_) return false;
  std::lock_guard<SpinMutex> lk(mu_);
  auto it = by_offset_.upper_bound(offset);
  if (it == by_offset_.begin()) return false;
  --it;
  const uint64_t off = it->first, len = it->second;
  if (offset < off || offset + need > off + len) return false;
  erase_free(it);
  if (offset > off) insert_free(off, offset - off);
  if (off + len > offset + need) insert_free(offset + need, off + len - offset - need);
  return true;
}

void PoolAllocator::free(const Range& range) {
  if (range.length == 0) return;
  std::lock_guard<SpinMutex> lk(mu_);
  uint64_t off = range.offset, len = range.length;
  auto next = by_offset_.lower_bound(off);
  if (next != by_offset_.begin()) {
    auto prev = std::prev(next);
    if (prev->first + prev->second == off) {
      off = prev->first;
      len += prev->second;
      erase_free(prev);
    } else if (prev->first + prev->second > off) {
      BB_LOG(ERROR) << "PoolAllocator(" << pool_id_ << "): double free / overlap at offset " << range.offset;
      return;
    }
  }
  next = by_offset_.lower_bound(range.offset);
  if (next != by_offset_.end()) {
    if (range.offset + range.length == next->first) {
      len += next->second;
      erase_free(next);
    } else if (range.offset + range.length > next->first) {
      BB_LOG(ERROR) << "PoolAllocator(" << pool_id_ << "): double free / overlap at offset " << range.offset;
      if (off != range.offset) insert_free(off, len - range.length);  // restore the merged predecessor
      return;
    }
  }
  insert_free(off, len);
}

size_t PoolAllocator::total_free() const { return free_pub_.load(std::memory_order_relaxed); }
size_t PoolAllocator::largest_free_block() const { return largest_pub_.load(std::memory_order_relaxed); }
double PoolAllocator::fragmentation_ratio() const {
  std::lock_guard<SpinMutex> lk(mu_);
  if (free_bytes_ == 0) return 0.0;
  return 1.0 - static_cast<double>(by_size_.rbegin()->first) / static_cast<double>(free_bytes_);
}
bool PoolAllocator::can_allocate(uint64_t size) const {
  const uint64_t need = aligned(size);
  std::lock_guard<SpinMutex> lk(mu_);
  return need == 0 || (!by_size_.empty() && by_size_.rbegin()->first >= need);
}
std::vector<Range> PoolAllocator::free_ranges() const {
  std::lock_guard<SpinMutex> lk(mu_);
  std::vector<Range> v;
  v.reserve(by_offset_.size());
  for (const auto& [o, l] : by_offset_) v.emplace_back(o, l);
  return v;
}
MemoryLocation PoolAllocator::to_memory_location(const Range& r) const {
  return MemoryLocation{base_addr_ + r.offset, rkey_, r.length};
}

// ================================================================ RangeAllocator
// Pool lookups happen ~20 times per put (ranking + shard construction).  Readers use a per-thread copy of the
// id -> allocator table that is refreshed only when the table's generation changes, so the hot path touches no
// shared lock word at all.  Retired allocators are parked in `graveyard_` (never freed while this object lives),
// which keeps every cached pointer valid.
namespace {
std::atomic<uint64_t> g_next_allocator_instance{1};
struct PoolTableCache {
  uint64_t instance = 0;
  uint64_t generation = 0;
  std::unordered_map<MemoryPoolId, PoolAllocator*> table;
};
thread_local PoolTableCache t_pool_cache;
}  // namespace

RangeAllocator::RangeAllocator() : instance_id_(g_next_allocator_instance.fetch_add(1)) {}

PoolAllocator* RangeAllocator::find_pool(const MemoryPoolId& id) const {
  PoolTableCache& c = t_pool_cache;
  const uint64_t gen = generation_.load(std::memory_order_acquire);
  if (c.instance != instance_id_ || c.generation != gen) {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    c.table.clear();
    for (const auto& [pid, pa] : pool_allocators_) c.table.emplace(pid, pa.get());
    c.instance = instance_id_;
    c.generation = generation_.load(std::memory_order_relaxed);
  }
  auto it = c.table.find(id);
  return it == c.table.end() ? nullptr : it->second;
}

PoolAllocator* RangeAllocator::ensure_pool(const MemoryPool& pool) {
  if (PoolAllocator* p = find_pool(pool.id)) return p;
  std::unique_lock<std::shared_mutex> lk(pools_mu_);
  auto& slot = pool_allocators_[pool.id];
  if (!slot) {
    // flash / disk pools are carved in 4 KiB blocks so a shard starts where O_DIRECT can reach it (the worker's own
    // allocator uses the same block for reserved extents); memory tiers keep the 256 B TMA granule
    slot = std::make_unique<PoolAllocator>(pool, is_disk_class(pool.storage_class) ? 4096 : PoolAllocator::kDefaultAlign);
    generation_.fetch_add(1, std::memory_order_release);
  }
  return slot.get();
}

std::vector<RangeAllocator::Candidate> RangeAllocator::rank_candidates(const AllocationRequest& req, const PoolMap& pools,
                                                                       bool* spill) const {
  std::vector<Candidate> v;
  v.reserve(pools.size());
  auto preferred = [&](StorageClass c) {
    return req.preferred_classes.empty() ||
           std::find(req.preferred_classes.begin(), req.preferred_classes.end(), c) != req.preferred_classes.end();
  };
  std::string client_domain;
  if (!req.client_node.empty())
    for (const auto& [id, p] : pools)
      if (p.node_id == req.client_node && !p.fabric_domain.empty()) client_domain = p.fabric_domain;
  for (const auto& [id, p] : pools) {
    if (!req.preferred_node.empty() && p.node_id != req.preferred_node) continue;
    if (std::find(req.exclude_pools.begin(), req.exclude_pools.end(), id) != req.exclude_pools.end()) continue;
    // NVLS replica arenas ("nvls-g<k>" domains) are reserved for symmetric multicast placement
    if (!req.symmetric_replicas && p.fabric_domain.compare(0, 5, "nvls-") == 0) continue;
    Candidate c;
    c.id = id;
    c.preferred = preferred(p.storage_class);
    c.locality = 0;
    if (req.enable_locality_awareness && !req.client_node.empty()) {
      if (p.node_id == req.client_node) c.locality = 2;
      else if (!client_domain.empty() && p.fabric_domain == client_domain) c.locality = 1;
    }
    c.bw = p.max_bw_gbps;
    if (const PoolAllocator* pa = find_pool(id)) {
      c.free_bytes = pa->total_free();
      c.largest = pa->largest_free_block();
    } else {
      c.free_bytes = p.size / PoolAllocator::kDefaultAlign * PoolAllocator::kDefaultAlign;
      c.largest = c.free_bytes;
    }
    c.worker = p.worker_id.empty() ? ("pool:" + id) : p.worker_id;
    v.push_back(std::move(c));
  }
  std::sort(v.begin(), v.end(), [](const Candidate& a, const Candidate& b) {
    if (a.preferred != b.preferred) return a.preferred;
    if (a.locality != b.locality) return a.locality > b.locality;
    if (a.bw != b.bw) return a.bw > b.bw;
    if (a.free_bytes != b.free_bytes) return a.free_bytes > b.free_bytes;
    return a.id < b.id;
  });
  if (spill) *spill = false;
  return v;
}

Result<ShardPlacement> RangeAllocator::make_shard(const MemoryPool& pool, const Range& r, uint64_t length) const {
  ShardPlacement s;
  s.pool_id = pool.id;
  s.worker_id = pool.worker_id;
  s.storage_class = pool.storage_class;
  s.length = length;
  const PoolAllocator* reg = find_pool(pool.id);
  if (reg) {
    if (!reg->registration_valid()) return ErrorCode::INVALID_PARAMETERS;
    s.endpoint = reg->endpoint();
  }
  switch (pool.storage_class) {
    case StorageClass::RAM_GPU:
      s.location = GpuSlabLocation{static_cast<uint32_t>(pool.gpu_device_id < 0 ? 0 : pool.gpu_device_id), 0, r.offset, length};
      break;
    case StorageClass::NVME:
    case StorageClass::SSD:
    case StorageClass::HDD:
      s.location = FileLocation{pool.mount_path.empty() ? pool.id : pool.mount_path + "/" + pool.id + ".dat", r.offset};
      break;
    case StorageClass::CXL_MEMORY:
    case StorageClass::CXL_TYPE2_DEVICE:
      s.location = CxlMemoryLocation{pool.id, r.offset / 256, r.offset, length};
      break;
    default: {
      const PoolAllocator* pa = find_pool(pool.id);
      MemoryLocation m = pa ? pa->to_memory_location(r) : MemoryLocation{pool.ucx_remote_addr + r.offset, 0, r.length};
      m.size = length;
      s.location = m;
    }
  }
  return s;
}

void RangeAllocator::rollback(const std::vector<Extent>& extents) {
  for (const auto& e : extents)
    if (PoolAllocator* pa = find_pool(e.pool)) pa->free(e.range);
}

bool RangeAllocator::single_shard(const AllocationRequest& req, size_t ncands) const {
  if (req.replication_factor != 1 || req.symmetric_replicas) return false;
  size_t wpc = std::min(std::max<size_t>(1, req.max_workers_per_copy), std::max<size_t>(1, ncands));  // as place() derives it
  if (!req.enable_striping || req.prefer_contiguous) wpc = 1;
  if (wpc > 1 && req.data_size / wpc < req.min_shard_size) {
    if (req.strict_min_shard) return false;
    wpc = std::max<size_t>(1, std::min(wpc, req.data_size / std::max<size_t>(1, req.min_shard_size)));
  }
  return wpc == 1;
}

Result<AllocationResult> RangeAllocator::place_single(const AllocationRequest& req, const PoolMap& pools, const std::vector<Candidate>& cands,
                                                      bool spill) {
  for (const Candidate& cd : cands) {
    const MemoryPool& pool = pools.at(cd.id);
    PoolAllocator* pa = ensure_pool(pool);
    auto r = pa->allocate(req.data_size, true);
    if (!r) continue;
    auto shard = make_shard(pool, *r, req.data_size);
    if (!shard.ok()) {
      pa->free(*r);
      return shard.error();
    }
    ObjectAllocation oa;
    oa.total_size = req.data_size;
    oa.extents.push_back({cd.id, *r, req.data_size});
    if (!ledger_insert(req.object_key, std::move(oa))) {
      pa->free(*r);
      return ErrorCode::OBJECT_ALREADY_EXISTS;
    }
    AllocationResult result;
    CopyPlacement copy;
    copy.copy_index = 0;
    copy.shards.push_back(std::move(shard.value()));
    result.copies.push_back(std::move(copy));
    result.total_shards_created = 1;
    result.pools_used = 1;
    result.stats.required_spillover = spill || !cd.preferred;
    result.stats.avg_shard_size = req.data_size;
    result.stats.fragmentation_score = static_cast<size_t>(pa->fragmentation_ratio() * 100.0);
    return result;
  }
  return ErrorCode::INSUFFICIENT_SPACE;
}

Result<AllocationResult> RangeAllocator::place(const AllocationRequest& req, const PoolMap& pools,
                                               const std::vector<Candidate>& cands, bool spill) {
  const size_t n = cands.size();
  const size_t repl = std::max<size_t>(1, req.replication_factor);
  size_t wpc = std::min(std::max<size_t>(1, req.max_workers_per_copy), n);
  if (!req.enable_striping || req.prefer_contiguous) wpc = 1;
  if (repl > 1 && n > wpc) {
    const size_t ideal = n / repl;  // leave room for the other replicas on distinct pools
    if (ideal >= 1) wpc = std::min(wpc, ideal);
  }
  if (wpc > 1 && req.data_size / wpc < req.min_shard_size) {
    if (req.strict_min_shard) return ErrorCode::INSUFFICIENT_SPACE;
    wpc = std::max<size_t>(1, std::min(wpc, req.data_size / std::max<size_t>(1, req.min_shard_size)));
  }

  AllocationResult result;
  std::vector<Extent> all;
  std::unordered_set<std::string> workers_used;  // failure domains taken by earlier copies
  std::unordered_set<MemoryPoolId> pools_used;
  size_t rot = 0;
  // Failure-domain partition: with >= repl distinct workers, worker w belongs to replica group
  // (rank of w) % repl, and copy c first tries to stay inside group c.  Losing one worker then
  // costs at most one copy (two pools of one worker never hold shards of different copies).
  std::unordered_map<std::string, size_t> worker_group;
  if (repl > 1) {
    std::vector<std::string> order;
    for (const auto& cd : cands)
      if (!worker_group.count(cd.worker)) {
        worker_group[cd.worker] = 0;
        order.push_back(cd.worker);
      }
    if (order.size() >= repl)
      for (size_t i = 0; i < order.size(); ++i) worker_group[order[i]] = i % repl;
    else
      worker_group.clear();
  }
  for (size_t c = 0; c < repl; ++c) {
    CopyPlacement copy;
    copy.copy_index = static_cast<uint32_t>(c);
    std::unordered_set<MemoryPoolId> in_copy;
    std::unordered_set<std::string> workers_in_copy;
    const size_t base = req.data_size / wpc, rem = req.data_size % wpc;
    for (size_t i = 0; i < wpc; ++i) {
      const uint64_t len = base + (i < rem ? 1 : 0);
      bool placed = false;
      // pass 0: a pool on a worker no other copy uses; pass 1: any pool not yet in this copy;
      // pass 2: anything with space (tiny clusters)
      for (int pass = 0; pass < 3 && !placed; ++pass) {
        for (size_t k = 0; k < n && !placed; ++k) {
          const Candidate& cd = cands[(rot + k) % n];
          if (pass < 2 && in_copy.count(cd.id)) continue;
          if (pass == 0) {
            if (!worker_group.empty() ? worker_group[cd.worker] != c : (workers_used.count(cd.worker) || pools_used.count(cd.id))) continue;
          }
          if (pass == 1 && repl > 1 && pools_used.count(cd.id) && n >= repl * wpc) continue;
          const MemoryPool& pool = pools.at(cd.id);
          PoolAllocator* pa = ensure_pool(pool);
          auto r = pa->allocate(len, true);
          if (!r) continue;
          auto shard = make_shard(pool, *r, len);
          if (!shard.ok()) {
            pa->free(*r);
            rollback(all);
            return shard.error();
          }
          all.push_back({cd.id, *r, len});
          copy.shards.push_back(std::move(shard.value()));
          in_copy.insert(cd.id);
          workers_in_copy.insert(cd.worker);
          if (!cd.preferred) spill = true;
          placed = true;
          rot = (rot + k + 1) % n;
        }
      }
      if (!placed) {
        rollback(all);
        return ErrorCode::INSUFFICIENT_SPACE;
      }
    }
    for (const auto& w : workers_in_copy) workers_used.insert(w);
    for (const auto& p : in_copy) pools_used.insert(p);
    result.total_shards_created += copy.shards.size();
    result.copies.push_back(std::move(copy));
  }

  // commit to the ledger
  {
    ObjectAllocation oa;
    oa.total_size = req.data_size;
    oa.extents = std::move(all);
    if (!ledger_insert(req.object_key, std::move(oa))) {
      rollback(oa.extents);
      return ErrorCode::OBJECT_ALREADY_EXISTS;
    }
  }
  result.pools_used = pools_used.size();
  result.stats.required_spillover = spill;
  result.stats.avg_shard_size = result.total_shards_created ? req.data_size * repl / result.total_shards_created : 0;
  double frag = 0;
  size_t cnt = 0;
  for (const auto& p : pools_used)
    if (PoolAllocator* pa = find_pool(p)) {
      frag += pa->fragmentation_ratio();
      ++cnt;
    }
  result.stats.fragmentation_score = cnt ? static_cast<size_t>(frag / static_cast<double>(cnt) * 100.0) : 0;
  return result;
}

Result<AllocationResult> RangeAllocator::place_symmetric(const AllocationRequest& req, const PoolMap& pools,
                                                         const std::vector<Candidate>& cands, bool spill) {
  const size_t repl = std::max<size_t>(1, req.replication_factor);
  // pick `repl` pools on distinct workers, best ranked first.  When the cluster advertises NVLS
  // replica arenas, all replicas must come from ONE multicast group (same "nvls-g<k>" domain):
  // the group with the most free space among those that can hold the object on `repl` members.
  std::unordered_map<std::string, std::vector<const Candidate*>> by_domain;
  for (const auto& c : cands) {
    const std::string& dom = pools.at(c.id).fabric_domain;
    if (dom.compare(0, 5, "nvls-") == 0 && c.largest >= req.data_size) by_domain[dom].push_back(&c);
  }
  std::vector<const Candidate*> chosen;
  if (!by_domain.empty()) {
    const std::vector<const Candidate*>* best = nullptr;
    uint64_t best_free = 0;
    std::string best_dom;
    for (const auto& [dom, v] : by_domain) {
      std::unordered_set<std::string> w;
      uint64_t min_free = ~0ull;
      for (auto* c : v) {
        w.insert(c->worker);
        min_free = std::min(min_free, c->free_bytes);
      }
      if (w.size() < repl) continue;
      if (!best || min_free > best_free || (min_free == best_free && dom < best_dom)) {
        best = &v;
        best_free = min_free;
        best_dom = dom;
      }
    }
    if (best) {
      std::unordered_set<std::string> workers;
      for (auto* c : *best) {
        if (workers.count(c->worker)) continue;
        chosen.push_back(c);
        workers.insert(c->worker);
        if (chosen.size() == repl) break;
      }
    }
  }
  if (chosen.size() < repl) {
    chosen.clear();
    std::unordered_set<std::string> workers;
    for (const auto& c : cands) {
      if (workers.count(c.worker)) continue;
      if (c.largest < req.data_size) continue;
      if (pools.at(c.id).fabric_domain.compare(0, 5, "nvls-") == 0) continue;  // partial groups cannot multicast
      chosen.push_back(&c);
      workers.insert(c.worker);
      if (chosen.size() == repl) break;
    }
  }
  if (chosen.size() < repl) return ErrorCode::INSUFFICIENT_SPACE;
  std::vector<PoolAllocator*> pas;
  for (auto* c : chosen) pas.push_back(ensure_pool(pools.at(c->id)));
  const uint64_t need = pas[0]->aligned(req.data_size);
  // Concurrent symmetric placements all compute the same "tightest common hole" and would take it from each other attempt
  // after attempt (8 ranks x 32 objects at once exhausted the retries on the 8-GPU box): they go one at a time.  What is left
  // to race with is ordinary placement on the same pools, which the retry loop absorbs.
  std::lock_guard<std::mutex> one_at_a_time(symmetric_mu_);
  for (int attempt = 0; attempt < 8; ++attempt) {
    // intersect the free lists (address ordered) and take the tightest common hole
    std::vector<Range> common = pas[0]->free_ranges();
    for (size_t i = 1; i < pas.size() && !common.empty(); ++i) {
      const std::vector<Range> other = pas[i]->free_ranges();
      std::vector<Range> out;
      size_t a = 0, b = 0;
      while (a < common.size() && b < other.size()) {
        const uint64_t lo = std::max(common[a].offset, other[b].offset);
        const uint64_t hi = std::min(common[a].end(), other[b].end());
        if (hi > lo) out.emplace_back(lo, hi - lo);
        if (common[a].end() < other[b].end()) ++a; else ++b;
      }
      common.swap(out);
    }
    const Range* best = nullptr;
    for (const auto& r : common)
      if (r.length >= need && (!best || r.length < best->length)) best = &r;
    if (!best) return ErrorCode::INSUFFICIENT_SPACE;
    const uint64_t off = best->offset;
    size_t got = 0;
    for (; got < pas.size(); ++got)
      if (!pas[got]->allocate_at(off, req.data_size)) break;
    if (got == pas.size()) {
      AllocationResult result;
      std::vector<Extent> all;
      for (size_t i = 0; i < pas.size(); ++i) {
        const MemoryPool& pool = pools.at(chosen[i]->id);
        Range r(off, need);
        auto shard = make_shard(pool, r, req.data_size);
        if (!shard.ok()) {
          for (auto* pa : pas) pa->free(r);
          return shard.error();
        }
        CopyPlacement cp;
        cp.copy_index = static_cast<uint32_t>(i);
        cp.shards.push_back(std::move(shard.value()));
        result.copies.push_back(std::move(cp));
        all.push_back({chosen[i]->id, r, req.data_size});
        if (!chosen[i]->preferred) spill = true;
      }
      ObjectAllocation oa;
      oa.total_size = req.data_size;
      oa.extents = std::move(all);
      if (!ledger_insert(req.object_key, std::move(oa))) {
        rollback(oa.extents);
        return ErrorCode::OBJECT_ALREADY_EXISTS;
      }
      result.total_shards_created = repl;
      result.pools_used = repl;
      result.stats.required_spillover = spill;
      result.stats.avg_shard_size = req.data_size;
      return result;
    }
    for (size_t i = 0; i < got; ++i) pas[i]->free(Range(off, need));  // raced with another writer: retry
  }
  return ErrorCode::ALLOCATION_FAILED;
}

Result<AllocationResult> RangeAllocator::allocate(const AllocationRequest& req, const PoolMap& pools) {
  if (req.object_key.empty()) return ErrorCode::INVALID_KEY;
  if (req.replication_factor == 0 || req.max_workers_per_copy == 0) return ErrorCode::INVALID_PARAMETERS;
  {
    LedgerShard& ls = ledger_for(req.object_key);
    std::lock_guard<SpinMutex> lk(ls.mu);
    if (ls.objects.count(req.object_key)) return ErrorCode::OBJECT_ALREADY_EXISTS;
  }
  if (pools.empty()) return ErrorCode::INSUFFICIENT_SPACE;
  bool spill = false;
  // Ranking walks every pool, copies ids and sorts: ~1 us that a batch of thousands of equal requests pays per object for
  // an answer that barely moves.  A thread keeps its last ranking and reuses it for up to kRankReuse consecutive requests
  // with the same shape against the same pool set (only the free-space figures are refreshed, from atomics; a placement that
  // fails on a reused ranking is retried on a fresh one).
  struct RankCache {
    const RangeAllocator* owner = nullptr;
    uint64_t gen = 0;
    size_t npools = 0;
    int left = 0;
    std::vector<StorageClass> classes;
    NodeId preferred_node, client_node;
    bool locality = false, symmetric = false;
    std::vector<Candidate> cands;
  };
  constexpr int kRankReuse = 256;
  thread_local RankCache rc;
  const bool cacheable = req.exclude_pools.empty();
  const uint64_t gen = generation_.load(std::memory_order_acquire);
  const bool hit = cacheable && rc.owner == this && rc.left > 0 && rc.gen == gen && rc.npools == pools.size() && rc.classes == req.preferred_classes &&
                   rc.preferred_node == req.preferred_node && rc.client_node == req.client_node && rc.locality == req.enable_locality_awareness &&
                   rc.symmetric == req.symmetric_replicas;
  if (hit) {
    --rc.left;
    // same pools, same shape: only the free-space figures move.  Refresh them from the allocators' live counters and
    // restore the order (the comparator is the one rank_candidates uses, so the outcome equals a fresh ranking)
    for (Candidate& c : rc.cands)
      if (const PoolAllocator* pa = find_pool(c.id)) {
        c.free_bytes = pa->total_free();
        c.largest = pa->largest_free_block();
      }
    std::sort(rc.cands.begin(), rc.cands.end(), [](const Candidate& a, const Candidate& b) {
      if (a.preferred != b.preferred) return a.preferred;
      if (a.locality != b.locality) return a.locality > b.locality;
      if (a.bw != b.bw) return a.bw > b.bw;
      if (a.free_bytes != b.free_bytes) return a.free_bytes > b.free_bytes;
      return a.id < b.id;
    });
    auto r = (req.symmetric_replicas && req.replication_factor > 1) ? place_symmetric(req, pools, rc.cands, spill)
             : single_shard(req, rc.cands.size())                       ? place_single(req, pools, rc.cands, spill)
                                                                        : place(req, pools, rc.cands, spill);
    if (r.ok()) return r;
    rc.left = 0;  // stale ranking (a pool filled up?): fall through to a fresh one
  }
  std::vector<Candidate> cands = rank_candidates(req, pools, &spill);
  if (cands.empty()) return ErrorCode::INSUFFICIENT_SPACE;
  auto r = (req.symmetric_replicas && req.replication_factor > 1) ? place_symmetric(req, pools, cands, spill)
           : single_shard(req, cands.size())                       ? place_single(req, pools, cands, spill)
                                                                   : place(req, pools, cands, spill);
  if (cacheable && r.ok()) {
    rc.owner = this;
    rc.gen = gen;
    rc.npools = pools.size();
    rc.left = kRankReuse;
    rc.classes = req.preferred_classes;
    rc.preferred_node = req.preferred_node;
    rc.client_node = req.client_node;
    rc.locality = req.enable_locality_awareness;
    rc.symmetric = req.symmetric_replicas;
    rc.cands = std::move(cands);
  }
  return r;
}

bool RangeAllocator::allocate_run(const AllocationRequest& shape, const std::vector<const ObjectKey*>& keys, const PoolMap& pools,
                                  std::vector<RunSlot>& out) {
  if (keys.empty() || shape.replication_factor != 1 || shape.max_workers_per_copy == 0 || shape.symmetric_replicas || !shape.exclude_pools.empty())
    return false;
  bool spill = false;
  std::vector<Candidate> cands = rank_candidates(shape, pools, &spill);
  {  // stripe width exactly as place() derives it; anything wider than one shard takes the general path
    size_t wpc = std::min(std::max<size_t>(1, shape.max_workers_per_copy), std::max<size_t>(1, cands.size()));
    if (!shape.enable_striping || shape.prefer_contiguous) wpc = 1;
    if (wpc > 1 && shape.data_size / wpc < shape.min_shard_size) {
      if (shape.strict_min_shard) return false;
      wpc = std::max<size_t>(1, std::min(wpc, shape.data_size / std::max<size_t>(1, shape.min_shard_size)));
    }
    if (wpc != 1) return false;
  }
  const size_t count = keys.size();
  out.assign(count, RunSlot{});
  if (cands.empty()) return true;  // every slot INSUFFICIENT_SPACE, as allocate() reports it
  size_t tie = 1;  // pools that rank equal but for their free space: the run is dealt out among them
  while (tie < cands.size() && cands[tie].preferred == cands[0].preferred && cands[tie].locality == cands[0].locality && cands[tie].bw == cands[0].bw) ++tie;
  struct Piece {
    const MemoryPool* pool;
    Range range;
  };
  std::vector<Piece> pieces(count, Piece{nullptr, Range(0, 0)});
  size_t next = 0;
  auto take = [&](const Candidate& cd, size_t want) {
    const MemoryPool& pool = pools.at(cd.id);
    PoolAllocator* pa = ensure_pool(pool);
    if (!pa->registration_valid()) return;
    const uint64_t need = pa->aligned(shape.data_size);
    int races = 0;
    while (want > 0 && next < count) {
      size_t chunk = want;
      if (need) {
        chunk = std::min<uint64_t>(want, pa->largest_free_block() / need);
        if (chunk == 0) return;
      }
      auto r = pa->allocate(static_cast<uint64_t>(chunk) * need, true);
      if (!r) {
        if (++races > 4) return;  // another writer took the hole between the probe and the call
        continue;
      }
      for (size_t k = 0; k < chunk; ++k) pieces[next + k] = Piece{&pool, Range(r->offset + k * need, need)};
      next += chunk;
      want -= chunk;
    }
  };
  const size_t per = (count + tie - 1) / tie;
  for (size_t c = 0; c < tie && next < count; ++c) take(cands[c], std::min(per, count - next));
  for (size_t c = 0; c < cands.size() && next < count; ++c) take(cands[c], count - next);
  // ledger + shard descriptors (a pool's used bytes are read off its allocator: nothing to account here)
  for (size_t i = 0; i < next; ++i) {
    const Piece& pc = pieces[i];
    const ObjectKey& key = *keys[i];
    auto shard = make_shard(*pc.pool, pc.range, shape.data_size);
    bool ok = shard.ok();
    if (ok) {
      ObjectAllocation oa;
      oa.total_size = shape.data_size;
      oa.extents.push_back({pc.pool->id, pc.range, shape.data_size});
      LedgerShard& ls = ledger_for(key);
      std::lock_guard<SpinMutex> lk(ls.mu);
      ok = ls.objects.emplace(key, std::move(oa)).second;
      if (!ok) out[i].status = ErrorCode::OBJECT_ALREADY_EXISTS;
    } else {
      out[i].status = shard.error();
    }
    if (!ok) {
      if (PoolAllocator* pa = find_pool(pc.pool->id)) pa->free(pc.range);
      continue;
    }
    out[i].status = ErrorCode::OK;
    out[i].shard = std::move(shard.value());
  }
  return true;
}

ErrorCode RangeAllocator::free(const ObjectKey& key) {
  ObjectAllocation oa;
  {
    LedgerShard& ls = ledger_for(key);
    std::lock_guard<SpinMutex> lk(ls.mu);
    auto it = ls.objects.find(key);
    if (it == ls.objects.end()) return ErrorCode::OBJECT_NOT_FOUND;
    oa = std::move(it->second);
    ls.objects.erase(it);
  }
  rollback(oa.extents);
  return ErrorCode::OK;
}

AllocatorStats RangeAllocator::get_stats(std::optional<StorageClass> sc) const {
  AllocatorStats st;
  size_t largest = 0;
  {
    std::shared_lock<std::shared_mutex> lk(pools_mu_);
    for (const auto& [id, pa] : pool_allocators_) {
      if (sc && pa->storage_class() != *sc) continue;
      st.total_free_bytes += pa->total_free();
      largest = std::max(largest, pa->largest_free_block());
      st.bytes_per_class[pa->storage_class()] += pa->capacity() - pa->total_free();
    }
  }
  for (const LedgerShard& ls : ledger_) {
    std::lock_guard<SpinMutex> lk(ls.mu);
    for (const auto& [key, oa] : ls.objects) {
      bool counted = false;
      for (const auto& e : oa.extents) {
        const PoolAllocator* pa = find_pool(e.pool);
        if (sc && (!pa || pa->storage_class() != *sc)) continue;
        st.total_allocated_bytes += e.range.length;
        ++st.total_shards;
        counted = true;
      }
      if (counted || (!sc && oa.extents.empty())) ++st.total_objects;
    }
  }
  st.fragmentation_ratio = st.total_free_bytes ? 1.0 - static_cast<double>(largest) / static_cast<double>(st.total_free_bytes) : 0.0;
  return st;
}

size_t RangeAllocator::get_free_space(StorageClass sc) const {
  size_t total = 0;
  std::shared_lock<std::shared_mutex> lk(pools_mu_);
  for (const auto& [id, pa] : pool_allocators_)
    if (pa->storage_class() == sc) total += pa->total_free();
  return total;
}

bool RangeAllocator::can_allocate(const AllocationRequest& req, const PoolMap& pools) const {
  if (req.replication_factor == 0 || req.max_workers_per_copy == 0 || pools.empty()) return false;
  bool spill = false;
  auto cands = rank_candidates(req, pools, &spill);
  if (cands.empty()) return false;
  // only preferred classes count (reference test "can_allocate class filter")
  uint64_t total = 0;
  size_t usable = 0;
  for (const auto& c : cands) {
    if (!c.preferred) continue;
    total += c.free_bytes;
    ++usable;
  }
  if (usable == 0) return false;
  const uint64_t need = static_cast<uint64_t>(req.data_size) * req.replication_factor;
  return total >= need;
}

void RangeAllocator::forget_pool(const MemoryPoolId& id) {
  {
    std::unique_lock<std::shared_mutex> lk(pools_mu_);
    auto it = pool_allocators_.find(id);
    if (it != pool_allocators_.end()) {
      graveyard_.push_back(std::move(it->second));  // other threads may still hold the pointer through their cache
      pool_allocators_.erase(it);
      generation_.fetch_add(1, std::memory_order_release);
    }
  }
  // The pool's memory is gone: drop its extents from every surviving object's ledger entry.  A worker that restarts
  // re-registers the same pool id with a fresh, all-free PoolAllocator; a stale extent left here would later be
  // "returned" to that new allocator by free() -> rollback() and hand a live object's range out a second time.
  for (LedgerShard& ls : ledger_) {
    std::lock_guard<SpinMutex> lk(ls.mu);
    for (auto& [key, oa] : ls.objects) {
      auto& ex = oa.extents;
      ex.erase(std::remove_if(ex.begin(), ex.end(), [&](const Extent& e) { return e.pool == id; }), ex.end());
    }
  }
}

// What is handed out on a pool is what its allocator does not have free: no second set of books next to the free lists
// (a global lock on every allocate and free, and one more thing to keep consistent across forget_pool / adopt / reset).
size_t RangeAllocator::pool_used_bytes(const MemoryPoolId& id) const {
  const PoolAllocator* pa = find_pool(id);
  return pa ? pa->used_bytes() : 0;
}

double RangeAllocator::pool_fragmentation(const MemoryPoolId& id) const {
  const PoolAllocator* pa = find_pool(id);
  return pa ? pa->fragmentation_ratio() : 0.0;
}

std::vector<ObjectKey> RangeAllocator::objects_on_pool(const MemoryPoolId& id) const {
  std::vector<ObjectKey> v;
  for (const LedgerShard& ls : ledger_) {
    std::lock_guard<SpinMutex> lk(ls.mu);
    for (const auto& [key, oa] : ls.objects)
      for (const auto& e : oa.extents)
        if (e.pool == id) {
          v.push_back(key);
          break;
        }
  }
  return v;
}

uint64_t pool_offset_of(const ShardPlacement& s, const MemoryPool& pool) {
  if (auto* g = std::get_if<GpuSlabLocation>(&s.location)) return g->offset;
  if (auto* f = std::get_if<FileLocation>(&s.location)) return f->file_offset;
  if (auto* x = std::get_if<CxlMemoryLocation>(&s.location)) return x->offset;
  if (auto* m = std::get_if<MemoryLocation>(&s.location)) {
    const uint64_t base = pool.ucx_remote_addr ? pool.ucx_remote_addr : pool.base_addr;
    return m->remote_addr >= base ? m->remote_addr - base : 0;
  }
  return 0;
}

size_t RangeAllocator::free_extents(const ObjectKey& key, const std::vector<ShardPlacement>& shards, const PoolMap& pools) {
  std::vector<Extent> gone;
  {
    LedgerShard& ls = ledger_for(key);
    std::lock_guard<SpinMutex> lk(ls.mu);
    auto it = ls.objects.find(key);
    if (it == ls.objects.end()) return 0;
    for (const auto& s : shards) {
      auto pit = pools.find(s.pool_id);
      if (pit == pools.end()) continue;
      const uint64_t off = pool_offset_of(s, pit->second);
      auto& ex = it->second.extents;
      auto e = std::find_if(ex.begin(), ex.end(), [&](const Extent& x) { return x.pool == s.pool_id && x.range.offset == off; });
      if (e == ex.end()) continue;
      it->second.total_size -= std::min<size_t>(it->second.total_size, e->length);
      gone.push_back(*e);
      ex.erase(e);
    }
    if (it->second.extents.empty()) ls.objects.erase(it);
  }
  rollback(gone);
  return gone.size();
}

ErrorCode RangeAllocator::adopt(const ObjectKey& key, const std::vector<CopyPlacement>& copies, const PoolMap& pools,
                                const MemoryPoolId& only_pool) {
  ObjectAllocation oa;
  for (const auto& c : copies)
    for (const auto& s : c.shards) {
      if (!only_pool.empty() && s.pool_id != only_pool) continue;
      auto pit = pools.find(s.pool_id);
      if (pit == pools.end()) continue;
      PoolAllocator* pa = ensure_pool(pit->second);
      const uint64_t off = pool_offset_of(s, pit->second);
      if (!pa->allocate_at(off, s.length)) {
        BB_LOG(WARNING) << "adopt: extent of " << key << " on " << s.pool_id << " is no longer free";
        continue;
      }
      oa.extents.push_back({s.pool_id, Range(off, pa->aligned(s.length)), s.length});
      oa.total_size += s.length;
    }
  if (only_pool.empty()) {
    if (!ledger_insert(key, std::move(oa))) {
      rollback(oa.extents);
      return ErrorCode::OBJECT_ALREADY_EXISTS;
    }
    return ErrorCode::OK;
  }
  // late adoption of one pool's extents: merge into the object's existing ledger entry (created by the first adopt)
  LedgerShard& ls = ledger_for(key);
  std::lock_guard<SpinMutex> lk(ls.mu);
  ObjectAllocation& dst = ls.objects[key];
  dst.total_size += oa.total_size;
  dst.extents.insert(dst.extents.end(), oa.extents.begin(), oa.extents.end());
  return ErrorCode::OK;
}

void RangeAllocator::reset() {
  {
    std::unique_lock<std::shared_mutex> lk(pools_mu_);
    for (auto& [id, pa] : pool_allocators_) graveyard_.push_back(std::move(pa));  // cached pointers stay valid
    pool_allocators_.clear();
    generation_.fetch_add(1, std::memory_order_release);
  }
  for (LedgerShard& ls : ledger_) {
    std::lock_guard<SpinMutex> lk(ls.mu);
    ls.objects.clear();
  }
}

bool RangeAllocator::ledger_insert(const ObjectKey& key, ObjectAllocation&& oa) {
  LedgerShard& ls = ledger_for(key);
  std::lock_guard<SpinMutex> lk(ls.mu);
  if (ls.objects.count(key)) return false;  // `oa` is left intact for the caller's rollback
  return ls.objects.emplace(key, std::move(oa)).second;
}

// ================================================================ factory / adapter
std::unique_ptr<IAllocator> AllocatorFactory::create(Strategy) { return std::make_unique<RangeAllocator>(); }
std::unique_ptr<IAllocator> AllocatorFactory::create_range_based() { return std::make_unique<RangeAllocator>(); }

KeystoneAllocatorAdapter::KeystoneAllocatorAdapter(std::unique_ptr<IAllocator> allocator) : allocator_(std::move(allocator)) {}

AllocationRequest KeystoneAllocatorAdapter::to_request(const ObjectKey& key, size_t data_size, const WorkerConfig& c) {
  AllocationRequest r;
  r.object_key = key;
  r.data_size = data_size;
  r.replication_factor = c.replication_factor;
  r.max_workers_per_copy = c.max_workers_per_copy;
  r.preferred_classes = c.preferred_classes;
  r.preferred_node = c.preferred_node;
  r.enable_locality_awareness = c.enable_locality_awareness;
  r.enable_striping = c.max_workers_per_copy > 1;
  r.prefer_contiguous = c.prefer_contiguous;
  r.min_shard_size = c.min_shard_size;
  r.symmetric_replicas = c.symmetric_replicas;
  return r;
}

Result<std::vector<CopyPlacement>> KeystoneAllocatorAdapter::allocate_data_copies(const ObjectKey& key, size_t data_size,
                                                                                  const WorkerConfig& config,
                                                                                  const IAllocator::PoolMap& pools,
                                                                                  const std::string& client_node,
                                                                                  const std::vector<MemoryPoolId>& exclude) {
  AllocationRequest r = to_request(key, data_size, config);
  r.client_node = client_node;
  r.exclude_pools = exclude;
  auto res = allocator_->allocate(r, pools);
  if (!res.ok()) return res.error();
  return std::move(res.value().copies);
}

bool KeystoneAllocatorAdapter::allocate_run(const std::vector<const ObjectKey*>& keys, size_t data_size, const WorkerConfig& config,
                                            const IAllocator::PoolMap& pools, const std::string& client_node,
                                            std::vector<IAllocator::RunSlot>& out) {
  AllocationRequest r = to_request(ObjectKey(), data_size, config);
  r.client_node = client_node;
  return allocator_->allocate_run(r, keys, pools, out);
}

ErrorCode KeystoneAllocatorAdapter::free_object(const ObjectKey& key) { return allocator_->free(key); }

AllocatorStats KeystoneAllocatorAdapter::get_allocator_stats(std::optional<StorageClass> sc) const { return allocator_->get_stats(sc); }

Result<bool> KeystoneAllocatorAdapter::can_allocate_object(size_t data_size, const WorkerConfig& config,
                                                           const IAllocator::PoolMap& pools) const {
  if (config.replication_factor == 0 || config.max_workers_per_copy == 0) return ErrorCode::INVALID_PARAMETERS;
  return allocator_->can_allocate(to_request("probe", data_size, config), pools);
}

}  // namespace bb::alloc
