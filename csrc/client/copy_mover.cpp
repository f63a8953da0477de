#include "client/copy_mover.h"

#include <cstring>
#include <map>
#include <mutex>

#include "common/log.h"
#include "net/tcp.h"
#include "rpc/wire.h"
#include "worker/worker_service.h"

namespace bb::client {

namespace {

uint64_t raw_offset(const ShardPlacement& s) {
  if (auto* g = std::get_if<GpuSlabLocation>(&s.location)) return g->offset;
  if (auto* f = std::get_if<FileLocation>(&s.location)) return f->file_offset;
  if (auto* c = std::get_if<CxlMemoryLocation>(&s.location)) return c->offset;
  if (auto* m = std::get_if<MemoryLocation>(&s.location)) return m->remote_addr | (1ull << 63);  // absolute
  return 0;
}

struct Conns {
  std::mutex mu;
  std::map<std::string, std::shared_ptr<net::RpcClient>> by_ep;
  int timeout_ms;
  std::shared_ptr<net::RpcClient> get(const std::string& ep) {
    std::lock_guard<std::mutex> lk(mu);
    auto& c = by_ep[ep];
    if (!c || !c->connected()) {
      auto hp = split_host_port(ep);
      if (!hp) return nullptr;
      c = std::make_shared<net::RpcClient>();
      if (c->connect(hp->first, static_cast<uint16_t>(hp->second), 3000) != ErrorCode::OK) {
        c.reset();
        return nullptr;
      }
    }
    return c;
  }
};

std::string ep_of(const ShardPlacement& s) { return s.endpoint.ip + ":" + std::to_string(s.endpoint.port); }

ErrorCode read_range(Conns& cs, const ShardPlacement& s, uint64_t off_in_shard, uint8_t* dst, uint64_t len) {
  auto c = cs.get(ep_of(s));
  if (!c) return ErrorCode::CONNECTION_FAILED;
  constexpr uint64_t kChunk = 8ull << 20;
  for (uint64_t pos = 0; pos < len; pos += kChunk) {
    const uint32_t n = static_cast<uint32_t>(std::min(kChunk, len - pos));
    wire::Writer w;
    w.str(s.pool_id);
    const uint64_t ro = raw_offset(s);
    w.u64((ro & (1ull << 63)) | ((ro & ~(1ull << 63)) + off_in_shard + pos));
    w.u32(n);
    auto r = c->call(worker::D_READ, w.data(), cs.timeout_ms);
    if (!r.ok() || r.value().size() < 4) return ErrorCode::TRANSFER_FAILED;
    uint32_t e;
    std::memcpy(&e, r.value().data(), 4);
    if (e != 0) return static_cast<ErrorCode>(e);
    if (r.value().size() != 4 + static_cast<size_t>(n)) return ErrorCode::TRANSFER_FAILED;
    std::memcpy(dst + pos, r.value().data() + 4, n);
  }
  return ErrorCode::OK;
}

ErrorCode write_range(Conns& cs, const ShardPlacement& s, const uint8_t* src, uint64_t len) {
  auto c = cs.get(ep_of(s));
  if (!c) return ErrorCode::CONNECTION_FAILED;
  constexpr uint64_t kChunk = 8ull << 20;
  for (uint64_t pos = 0; pos < len; pos += kChunk) {
    const uint32_t n = static_cast<uint32_t>(std::min(kChunk, len - pos));
    wire::Writer w;
    w.str(s.pool_id);
    const uint64_t ro = raw_offset(s);
    w.u64((ro & (1ull << 63)) | ((ro & ~(1ull << 63)) + pos));
    w.u32(n);
    w.raw(src + pos, n);
    auto r = c->call(worker::D_WRITE, w.data(), cs.timeout_ms);
    if (!r.ok()) return ErrorCode::TRANSFER_FAILED;
    wire::Reader rd(r.value());
    const ErrorCode ec = rd.ec();
    if (ec != ErrorCode::OK) return ec;
  }
  return ErrorCode::OK;
}

}  // namespace

keystone::CopyMover make_data_server_mover(size_t io_parallelism, int rpc_timeout_ms) {
  auto conns = std::make_shared<Conns>();
  conns->timeout_ms = rpc_timeout_ms;
  (void)io_parallelism;
  return [conns](const ObjectKey& key, const CopyPlacement& src, CopyPlacement& dst, ChecksumAlgo algo) -> ErrorCode {
    // fast path: identical shard layout and both shards served by the same worker -> D_COPY
    bool same_layout = src.shards.size() == dst.shards.size();
    for (size_t i = 0; same_layout && i < src.shards.size(); ++i)
      same_layout = src.shards[i].length == dst.shards[i].length && ep_of(src.shards[i]) == ep_of(dst.shards[i]);
    if (same_layout) {
      for (size_t i = 0; i < src.shards.size(); ++i) {
        auto c = conns->get(ep_of(src.shards[i]));
        if (!c) return ErrorCode::CONNECTION_FAILED;
        wire::Writer w;
        w.str(src.shards[i].pool_id);
        w.u64(raw_offset(src.shards[i]));
        w.str(dst.shards[i].pool_id);
        w.u64(raw_offset(dst.shards[i]));
        w.u64(src.shards[i].length);
        w.u32(static_cast<uint32_t>(algo));
        auto r = c->call(worker::D_COPY, w.data(), conns->timeout_ms);
        if (!r.ok()) return ErrorCode::TRANSFER_FAILED;
        wire::Reader rd(r.value());
        const ErrorCode ec = rd.ec();
        if (ec != ErrorCode::OK) return ec;
        const uint64_t digest = rd.u64();
        if (algo != ChecksumAlgo::NONE && src.shards[i].checksum_algo == algo && src.shards[i].checksum != digest) {
          BB_LOG(ERROR) << "mover: source shard of " << key << " is corrupt (digest mismatch)";
          return ErrorCode::CHECKSUM_MISMATCH;
        }
        dst.shards[i].checksum = digest;
        dst.shards[i].checksum_algo = algo;
      }
      return ErrorCode::OK;
    }
    // GPU tier on both sides, different workers: the destination worker pulls the shard out of the source slab
    // itself (CUDA IPC mapping + one fused-kernel launch over NVLink); nothing passes through this process.
    bool gpu_pull = src.shards.size() == dst.shards.size() && !src.shards.empty();
    for (size_t i = 0; gpu_pull && i < src.shards.size(); ++i)
      gpu_pull = src.shards[i].length == dst.shards[i].length && src.shards[i].storage_class == StorageClass::RAM_GPU &&
                 dst.shards[i].storage_class == StorageClass::RAM_GPU && !src.shards[i].endpoint.worker_key.empty();
    if (gpu_pull) {
      bool fell_back = false;
      for (size_t i = 0; i < src.shards.size() && !fell_back; ++i) {
        auto c = conns->get(ep_of(dst.shards[i]));
        if (!c) return ErrorCode::CONNECTION_FAILED;
        wire::Writer w;
        w.str(dst.shards[i].pool_id);
        w.u64(raw_offset(dst.shards[i]));
        const auto& k = src.shards[i].endpoint.worker_key;
        w.str(std::string(k.begin(), k.end()));
        w.u64(raw_offset(src.shards[i]));
        w.u64(src.shards[i].length);
        w.u32(static_cast<uint32_t>(algo));
        auto r = c->call(worker::D_PULL, w.data(), conns->timeout_ms);
        if (!r.ok()) return ErrorCode::TRANSFER_FAILED;
        wire::Reader rd(r.value());
        const ErrorCode ec = rd.ec();
        if (ec == ErrorCode::NOT_IMPLEMENTED && i == 0) {  // e.g. both slabs in one process: relay below
          fell_back = true;
          break;
        }
        if (ec != ErrorCode::OK) return ec;
        const uint64_t digest = rd.u64();
        if (algo != ChecksumAlgo::NONE && src.shards[i].checksum_algo == algo && src.shards[i].checksum != digest) {
          BB_LOG(ERROR) << "mover: source shard of " << key << " is corrupt (digest mismatch)";
          return ErrorCode::CHECKSUM_MISMATCH;
        }
        dst.shards[i].checksum = digest;
        dst.shards[i].checksum_algo = algo;
      }
      if (!fell_back) return ErrorCode::OK;
    }
    // general path: relay the object through this process
    uint64_t total = 0;
    for (const auto& s : src.shards) total += s.length;
    uint64_t dtotal = 0;
    for (const auto& s : dst.shards) dtotal += s.length;
    if (total != dtotal) return ErrorCode::INVALID_PARAMETERS;
    std::vector<uint8_t> buf(total);
    uint64_t off = 0;
    for (const auto& s : src.shards) {
      ErrorCode ec = read_range(*conns, s, 0, buf.data() + off, s.length);
      if (ec != ErrorCode::OK) return ec;
      if (algo != ChecksumAlgo::NONE && s.checksum_algo == algo && checksum(algo, buf.data() + off, s.length) != s.checksum) {
        BB_LOG(ERROR) << "mover: source shard of " << key << " is corrupt (digest mismatch)";
        return ErrorCode::CHECKSUM_MISMATCH;
      }
      off += s.length;
    }
    off = 0;
    for (auto& s : dst.shards) {
      ErrorCode ec = write_range(*conns, s, buf.data() + off, s.length);
      if (ec != ErrorCode::OK) return ec;
      s.checksum = checksum(algo, buf.data() + off, s.length);
      s.checksum_algo = algo;
      off += s.length;
    }
    return ErrorCode::OK;
  };
}

keystone::CopyVerifier make_data_server_verifier(int rpc_timeout_ms) {
  auto conns = std::make_shared<Conns>();
  conns->timeout_ms = rpc_timeout_ms;
  return [conns](const ObjectKey& key, const CopyPlacement& copy, ChecksumAlgo algo) -> ErrorCode {
    for (const auto& s : copy.shards) {
      if (algo == ChecksumAlgo::NONE || s.checksum_algo != algo || s.length == 0) continue;  // nothing recorded to compare with
      auto c = conns->get(ep_of(s));
      if (!c) return ErrorCode::CONNECTION_FAILED;
      wire::Writer w;
      w.str(s.pool_id);
      w.u64(raw_offset(s));
      w.u64(s.length);
      w.u32(static_cast<uint32_t>(algo));
      auto r = c->call(worker::D_CHECKSUM, w.data(), conns->timeout_ms);
      if (!r.ok()) return ErrorCode::TRANSFER_FAILED;
      wire::Reader rd(r.value());
      const ErrorCode ec = rd.ec();
      if (ec != ErrorCode::OK) return ec;
      const uint64_t digest = rd.u64();
      if (!rd.ok()) return ErrorCode::TRANSFER_FAILED;
      if (digest != s.checksum) {
        BB_LOG(ERROR) << "scrub: shard of " << key << " on " << s.pool_id << " does not match its digest";
        return ErrorCode::CHECKSUM_MISMATCH;
      }
    }
    return ErrorCode::OK;
  };
}

keystone::ReservationHooks make_data_server_reservation_hooks(int rpc_timeout_ms) {
  auto conns = std::make_shared<Conns>();
  conns->timeout_ms = rpc_timeout_ms;
  struct Group {
    std::string ep, pool;
    std::vector<std::pair<size_t, size_t>> at;  // (copy, shard)
  };
  auto group = [](const std::vector<CopyPlacement>& copies) {
    std::map<std::pair<std::string, std::string>, Group> g;
    for (size_t c = 0; c < copies.size(); ++c)
      for (size_t s = 0; s < copies[c].shards.size(); ++s) {
        const auto& sh = copies[c].shards[s];
        Group& x = g[{ep_of(sh), sh.pool_id}];
        x.ep = ep_of(sh);
        x.pool = sh.pool_id;
        x.at.emplace_back(c, s);
      }
    return g;
  };
  auto call_ec = [conns](const std::string& ep, uint32_t method, const std::string& req, std::string* body = nullptr) {
    auto c = conns->get(ep);
    if (!c) return ErrorCode::CONNECTION_FAILED;
    auto r = c->call(method, req, conns->timeout_ms);
    if (!r.ok()) return ErrorCode::TRANSFER_FAILED;
    wire::Reader rd(r.value());
    const ErrorCode ec = rd.ec();
    if (body && ec == ErrorCode::OK) body->assign(r.value(), 4, std::string::npos);
    return ec;
  };
  keystone::ReservationHooks h;
  auto by_tokens = [group, call_ec](uint32_t method) {
    return [group, call_ec, method](const std::vector<CopyPlacement>& copies, const keystone::ShardTokens& tokens) {
      ErrorCode worst = ErrorCode::OK;
      for (const auto& [k, g] : group(copies)) {
        wire::Writer w;
        w.str(g.pool);
        uint32_t n = 0;
        for (const auto& [c, s] : g.at) n += (c < tokens.size() && s < tokens[c].size() && !tokens[c][s].empty()) ? 1u : 0u;
        if (!n) continue;
        w.u32(n);
        for (const auto& [c, s] : g.at)
          if (c < tokens.size() && s < tokens[c].size() && !tokens[c][s].empty()) w.str(tokens[c][s]);
        const ErrorCode ec = call_ec(g.ep, method, w.data());
        if (ec != ErrorCode::OK && worst == ErrorCode::OK) worst = ec;
      }
      return worst;
    };
  };
  h.commit = by_tokens(worker::D_COMMIT);
  h.abort = by_tokens(worker::D_ABORT);
  h.reserve = [group, call_ec, abort = h.abort](const ObjectKey& key, const std::vector<CopyPlacement>& copies, uint64_t ttl_ms,
                                                keystone::ShardTokens* tokens) {
    tokens->assign(copies.size(), {});
    for (size_t c = 0; c < copies.size(); ++c) (*tokens)[c].assign(copies[c].shards.size(), std::string());
    ErrorCode ec = ErrorCode::OK;
    for (const auto& [k, g] : group(copies)) {
      wire::Writer w;
      w.str(g.pool);
      w.str(key);
      w.u64(ttl_ms);
      w.u32(static_cast<uint32_t>(g.at.size()));
      for (const auto& [c, s] : g.at) {
        w.u64(raw_offset(copies[c].shards[s]));
        w.u64(copies[c].shards[s].length);
      }
      std::string body;
      ec = call_ec(g.ep, worker::D_RESERVE, w.data(), &body);
      if (ec != ErrorCode::OK) break;
      wire::Reader rd(body);
      const uint32_t n = rd.u32();
      if (n != g.at.size()) {
        ec = ErrorCode::TRANSFER_FAILED;
        break;
      }
      for (const auto& [c, s] : g.at) (*tokens)[c][s] = rd.str();
    }
    if (ec != ErrorCode::OK) abort(copies, *tokens);  // all or nothing across workers
    return ec;
  };
  h.release = [group, call_ec](const std::vector<CopyPlacement>& copies) {
    ErrorCode worst = ErrorCode::OK;
    for (const auto& [k, g] : group(copies)) {
      wire::Writer w;
      w.str(g.pool);
      w.u32(static_cast<uint32_t>(g.at.size()));
      for (const auto& [c, s] : g.at) {
        w.u64(raw_offset(copies[c].shards[s]));
        w.u64(copies[c].shards[s].length);
      }
      const ErrorCode ec = call_ec(g.ep, worker::D_FREE, w.data());
      if (ec != ErrorCode::OK && worst == ErrorCode::OK) worst = ec;
    }
    return worst;
  };
  return h;
}

}  // namespace bb::client
