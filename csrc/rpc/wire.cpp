#include "rpc/wire.h"

namespace bb::wire {

void put(Writer& w, const TransportEndpoint& e) {
  w.str(e.ip);
  w.u32(e.port);
  w.bytes(e.worker_key);
}
void get(Reader& r, TransportEndpoint& e) {
  e.ip = r.str();
  e.port = r.u32();
  e.worker_key = r.bytes();
}

void put(Writer& w, const LocationDetail& l) {
  w.u8(static_cast<uint8_t>(l.index()));
  if (auto* m = std::get_if<MemoryLocation>(&l)) {
    w.u64(m->remote_addr);
    w.u32(m->rkey);
    w.u64(m->size);
  } else if (auto* f = std::get_if<FileLocation>(&l)) {
    w.str(f->file_path);
    w.u64(f->file_offset);
  } else if (auto* c = std::get_if<CxlMemoryLocation>(&l)) {
    w.str(c->device_id);
    w.u64(c->region_id);
    w.u64(c->offset);
    w.u64(c->size);
  } else if (auto* g = std::get_if<GpuSlabLocation>(&l)) {
    w.u32(g->device_rank);
    w.u32(g->slab_id);
    w.u64(g->offset);
    w.u64(g->size);
  }
}
void get(Reader& r, LocationDetail& l) {
  switch (r.u8()) {
    case 0: {
      MemoryLocation m;
      m.remote_addr = r.u64();
      m.rkey = r.u32();
      m.size = r.u64();
      l = m;
      break;
    }
    case 1: {
      FileLocation f;
      f.file_path = r.str();
      f.file_offset = r.u64();
      l = f;
      break;
    }
    case 2: {
      CxlMemoryLocation c;
      c.device_id = r.str();
      c.region_id = r.u64();
      c.offset = r.u64();
      c.size = r.u64();
      l = c;
      break;
    }
    case 3: {
      GpuSlabLocation g;
      g.device_rank = r.u32();
      g.slab_id = r.u32();
      g.offset = r.u64();
      g.size = r.u64();
      l = g;
      break;
    }
    default:
      r.fail();  // unknown variant tag
  }
}

void put(Writer& w, const ShardPlacement& s) {
  w.str(s.pool_id);
  w.str(s.worker_id);
  put(w, s.endpoint);
  w.u32(static_cast<uint32_t>(s.storage_class));
  w.u64(s.length);
  put(w, s.location);
  w.u64(s.checksum);
  w.u32(static_cast<uint32_t>(s.checksum_algo));
}
void get(Reader& r, ShardPlacement& s) {
  s.pool_id = r.str();
  s.worker_id = r.str();
  get(r, s.endpoint);
  s.storage_class = static_cast<StorageClass>(r.u32());
  s.length = r.u64();
  get(r, s.location);
  s.checksum = r.u64();
  s.checksum_algo = static_cast<ChecksumAlgo>(r.u32());
}

void put(Writer& w, const CopyPlacement& c) {
  w.u32(c.copy_index);
  w.u32(static_cast<uint32_t>(c.shards.size()));
  for (const auto& s : c.shards) put(w, s);
}
void get(Reader& r, CopyPlacement& c) {
  c.copy_index = r.u32();
  const uint32_t n = r.count(16);
  c.shards.resize(n);
  for (auto& s : c.shards) get(r, s);
}

void put(Writer& w, const std::vector<CopyPlacement>& v) {
  w.u32(static_cast<uint32_t>(v.size()));
  for (const auto& c : v) put(w, c);
}
void get(Reader& r, std::vector<CopyPlacement>& v) {
  const uint32_t n = r.count(8);
  v.resize(n);
  for (auto& c : v) get(r, c);
}

namespace {
// Where a shard sits inside its pool -- the one location field two shards of one pool differ in.
uint64_t position_of(const LocationDetail& l) {
  if (auto* m = std::get_if<MemoryLocation>(&l)) return m->remote_addr;
  if (auto* f = std::get_if<FileLocation>(&l)) return f->file_offset;
  if (auto* c = std::get_if<CxlMemoryLocation>(&l)) return c->offset;
  return std::get<GpuSlabLocation>(l).offset;
}
void set_position(LocationDetail& l, uint64_t pos) {
  if (auto* m = std::get_if<MemoryLocation>(&l)) m->remote_addr = pos;
  else if (auto* f = std::get_if<FileLocation>(&l)) f->file_offset = pos;
  else if (auto* c = std::get_if<CxlMemoryLocation>(&l)) c->offset = pos, c->region_id = pos / 256;
  else std::get<GpuSlabLocation>(l).offset = pos;
}
bool same_but_position(const LocationDetail& a, const LocationDetail& b) {
  if (a.index() != b.index()) return false;
  if (auto* m = std::get_if<MemoryLocation>(&a)) {
    const auto& n = std::get<MemoryLocation>(b);
    return m->rkey == n.rkey && m->size == n.size;
  }
  if (auto* f = std::get_if<FileLocation>(&a)) return f->file_path == std::get<FileLocation>(b).file_path;
  if (auto* c = std::get_if<CxlMemoryLocation>(&a)) {
    const auto& d = std::get<CxlMemoryLocation>(b);
    return c->device_id == d.device_id && c->size == d.size && c->region_id == c->offset / 256 && d.region_id == d.offset / 256;
  }
  const auto& g = std::get<GpuSlabLocation>(a);
  const auto& h = std::get<GpuSlabLocation>(b);
  return g.device_rank == h.device_rank && g.slab_id == h.slab_id && g.size == h.size;
}
constexpr uint8_t kPlacementFull = 0, kPlacementDelta = 1;
}  // namespace

void PlacementBatchWriter::put(const Result<std::vector<CopyPlacement>>& res) {
  w_.ec(res.error());
  if (!res.ok()) return;
  const auto& v = res.value();
  const ShardPlacement* s = (v.size() == 1 && v[0].copy_index == 0 && v[0].shards.size() == 1) ? &v[0].shards[0] : nullptr;
  if (s && base_ && s->length == base_->length && s->storage_class == base_->storage_class && s->checksum_algo == base_->checksum_algo &&
      s->pool_id == base_->pool_id && s->worker_id == base_->worker_id && s->endpoint == base_->endpoint &&
      same_but_position(s->location, base_->location)) {
    w_.u8(kPlacementDelta);
    w_.u64(position_of(s->location));
    w_.u64(s->checksum);
    return;
  }
  w_.u8(kPlacementFull);
  wire::put(w_, v);
  base_ = s;  // results live until the reply is written
}

Result<std::vector<CopyPlacement>> PlacementBatchReader::get() {
  const ErrorCode ec = r_.ec();
  if (ec != ErrorCode::OK) return ec;
  const uint8_t tag = r_.u8();
  if (tag == kPlacementDelta) {
    const uint64_t pos = r_.u64(), sum = r_.u64();
    if (!r_.ok() || !have_base_) {
      r_.fail();
      return ErrorCode::RPC_FAILED;
    }
    std::vector<CopyPlacement> v = base_;
    set_position(v[0].shards[0].location, pos);
    v[0].shards[0].checksum = sum;
    return v;
  }
  std::vector<CopyPlacement> v;
  if (tag == kPlacementFull) wire::get(r_, v);
  else r_.fail();
  if (!r_.ok()) return ErrorCode::RPC_FAILED;
  have_base_ = v.size() == 1 && v[0].copy_index == 0 && v[0].shards.size() == 1;
  if (have_base_) base_ = v;
  return v;
}

void put(Writer& w, const WorkerConfig& c) {
  w.u64(c.replication_factor);
  w.u64(c.max_workers_per_copy);
  w.boolean(c.enable_soft_pin);
  w.str(c.preferred_node);
  w.u32(static_cast<uint32_t>(c.preferred_classes.size()));
  for (auto sc : c.preferred_classes) w.u32(static_cast<uint32_t>(sc));
  w.u64(c.ttl_ms);
  w.boolean(c.enable_locality_awareness);
  w.boolean(c.prefer_contiguous);
  w.u64(c.min_shard_size);
  w.u32(static_cast<uint32_t>(c.checksum));
  w.boolean(c.pack_fp8);
  w.boolean(c.symmetric_replicas);
}
void get(Reader& r, WorkerConfig& c) {
  c.replication_factor = r.u64();
  c.max_workers_per_copy = r.u64();
  c.enable_soft_pin = r.boolean();
  c.preferred_node = r.str();
  const uint32_t n = r.count(4);
  c.preferred_classes.resize(n);
  for (auto& sc : c.preferred_classes) sc = static_cast<StorageClass>(r.u32());
  c.ttl_ms = r.u64();
  c.enable_locality_awareness = r.boolean();
  c.prefer_contiguous = r.boolean();
  c.min_shard_size = r.u64();
  c.checksum = static_cast<ChecksumAlgo>(r.u32());
  c.pack_fp8 = r.boolean();
  c.symmetric_replicas = r.boolean();
}

void put(Writer& w, const ClusterStats& s) {
  w.u64(s.total_workers);
  w.u64(s.total_memory_pools);
  w.u64(s.total_objects);
  w.u64(s.total_capacity);
  w.u64(s.used_capacity);
  w.f64(s.avg_utilization);
  w.u64(s.pending_objects);
  w.u64(s.active_clients);
}
void get(Reader& r, ClusterStats& s) {
  s.total_workers = r.u64();
  s.total_memory_pools = r.u64();
  s.total_objects = r.u64();
  s.total_capacity = r.u64();
  s.used_capacity = r.u64();
  s.avg_utilization = r.f64();
  s.pending_objects = r.u64();
  s.active_clients = r.u64();
}

void put(Writer& w, const MemoryPool& p) {
  w.str(p.id);
  w.str(p.node_id);
  w.str(p.worker_id);
  w.u64(p.base_addr);
  w.u64(p.size);
  w.u64(p.used);
  w.u32(static_cast<uint32_t>(p.storage_class));
  w.str(p.ucx_endpoint);
  w.u64(p.ucx_remote_addr);
  w.str(p.ucx_rkey_hex);
  w.i64(p.gpu_device_id);
  w.i64(p.numa_node);
  w.f64(p.max_bw_gbps);
  w.str(p.fabric_domain);
  w.str(p.mount_path);
}
void get(Reader& r, MemoryPool& p) {
  p.id = r.str();
  p.node_id = r.str();
  p.worker_id = r.str();
  p.base_addr = r.u64();
  p.size = r.u64();
  p.used = r.u64();
  p.storage_class = static_cast<StorageClass>(r.u32());
  p.ucx_endpoint = r.str();
  p.ucx_remote_addr = r.u64();
  p.ucx_rkey_hex = r.str();
  p.gpu_device_id = static_cast<int32_t>(r.i64());
  p.numa_node = static_cast<int32_t>(r.i64());
  p.max_bw_gbps = r.f64();
  p.fabric_domain = r.str();
  p.mount_path = r.str();
}

}  // namespace bb::wire
