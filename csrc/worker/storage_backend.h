// Storage tiers of a worker.
//
// Parity: reference include/blackbird/worker/storage/storage_backend.h:14-133 — two-phase shard
// lifecycle `reserve_shard -> commit_shard | abort_shard`, later `free_shard(addr, size)`;
// ReservationToken (:14-25), StorageStats (:30-41), factory (:131-133).  Implementations:
// RamBackend (ram_backend.h:13-69), MmapDiskBackend (mmap_disk_backend.h:22-101),
// IoUringDiskBackend (iouring_disk_backend.h:21-127), CxlMemoryBackend (cxl_memory_backend.h:13-112).
//
// Differences by design:
//  * every backend has a real data API (`write` / `read` at pool offsets, plus `direct_ptr` for
//    memory-mapped tiers) — the reference backends can only hand out addresses;
//  * reservations are carved from an embedded PoolAllocator, so uncommitted reservations can
//    never overlap (reference bug §2.8 #11) and there is no recursive locking (#10);
//  * the io_uring tier really submits SQEs (READ/WRITE with registered, aligned staging
//    buffers) through raw syscalls — liburing is not needed;
//  * the disk tiers keep a manifest with per-extent CRC32C so a restarted worker can
//    re-advertise what it holds (§5.4);
//  * the factory builds the disk tiers (reference returns nullptr, ram_backend.cpp:299-301);
//  * a GPU tier exists (fabric/gpu_slab.h) — `RAM_GPU` is `malloc` in the reference.
#pragma once
#include <atomic>
#include <chrono>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "alloc/allocator.h"
#include "common/types.h"
#include "net/aead.h"

namespace bb::worker {

struct ReservationToken {
  std::string token_id;
  MemoryPoolId pool_id;
  uint64_t remote_addr = 0;  // base_address + offset
  uint64_t rkey = 0;
  uint64_t size = 0;
  std::chrono::system_clock::time_point expires_at;
};

struct StorageStats {
  uint64_t total_capacity = 0;
  uint64_t used_capacity = 0;       // committed + reserved
  uint64_t available_capacity = 0;
  uint64_t num_reservations = 0;
  uint64_t num_committed_shards = 0;
  double utilization = 0.0;
  double fragmentation = 0.0;
  uint64_t bytes_written = 0;
  uint64_t bytes_read = 0;
  uint64_t io_errors = 0;
};

struct BackendOptions {
  std::string mount_path;          // disk tiers / dax device path
  int gpu_device_id = 0;           // GPU tier
  int numa_node = -1;              // CXL / DRAM binding
  uint32_t queue_depth = 64;       // io_uring
  bool pin_memory = false;         // DRAM tier: cudaHostRegister when CUDA is present
  // DRAM tier: back the pool with a memfd and advertise "file:/proc/<pid>/fd/<n>" as its registration key, so that
  // GPU clients in OTHER processes of the same host can map it (and register it with their CUDA context): the fused
  // kernels then read / write the DRAM tier directly over PCIe instead of going through the TCP data server.
  bool shared_memory = false;
  uint64_t reservation_ttl_ms = 10 * 60 * 1000;
  uint64_t interleave_granularity = 256;  // CXL region id granularity
  bool persistent = false;                // CXL persistent mode: msync on commit/flush
  // File-backed tiers (NVMe / SSD / HDD): what reaches the file is AES-256-CTR of the pool bytes, keyed from this
  // passphrase (worker `at_rest_key:` / BB_AT_REST_KEY) and the pool id.  Empty = plain.  (net/aead.h OffsetCipher)
  std::string at_rest_key;
  std::string at_rest_scope;  // what the key and counter prefix are bound to besides the passphrase: the pool id, known to the
                              // worker before the backend exists (initialize() may already read extents back)
};

class StorageBackend {
 public:
  virtual ~StorageBackend() = default;
  virtual StorageClass get_storage_class() const = 0;
  virtual uint64_t get_total_capacity() const = 0;
  virtual uint64_t get_used_capacity() const;
  virtual uint64_t get_available_capacity() const;
  virtual uint64_t get_base_address() const = 0;  // address clients add offsets to
  virtual uint64_t get_rkey() const = 0;

  virtual ErrorCode initialize() = 0;
  virtual void shutdown() = 0;

  // ---- two-phase shard lifecycle
  virtual Result<ReservationToken> reserve_shard(uint64_t size, const std::string& hint = "");
  virtual ErrorCode commit_shard(const ReservationToken& token);
  virtual ErrorCode abort_shard(const ReservationToken& token);
  virtual ErrorCode free_shard(uint64_t remote_addr, uint64_t size);
  // Reservation protocol as the Keystone drives it (it owns the placement decision, so the extent is given): reserves
  // exactly [offset, offset + size) for `owner` (the object key) until `ttl_ms` from now (0 = the backend's default).
  // ALLOCATION_FAILED when any part of the range is already reserved or committed.
  virtual Result<ReservationToken> reserve_shard_at(uint64_t offset, uint64_t size, const std::string& owner, uint64_t ttl_ms = 0);
  ErrorCode commit_shard_id(const std::string& token_id);
  ErrorCode abort_shard_id(const std::string& token_id);
  // Reclaims every reservation whose token has expired (the writer vanished between put_start and put_complete) and
  // returns (token id, owner) of each, so that the worker can tell the Keystone.  This is what the reference's
  // `allocation_poll_interval_ms` (include/blackbird/worker/worker_service.h:37) was meant to drive.
  std::vector<std::pair<std::string, std::string>> reap_expired_reservations();
  virtual StorageStats get_stats() const;

  // ---- data plane (offsets are relative to the pool base)
  virtual ErrorCode write(uint64_t offset, const void* data, uint64_t len) = 0;
  virtual ErrorCode read(uint64_t offset, void* data, uint64_t len) = 0;
  virtual void* direct_ptr(uint64_t /*offset*/) { return nullptr; }  // host-mapped tiers only
  virtual ErrorCode flush() { return ErrorCode::OK; }
  // True when a CUDA kernel in this process can address direct_ptr() (device memory, or host memory
  // registered with the driver): such tiers take part in fused-kernel tier moves.
  virtual bool cuda_accessible() const { return false; }
  // Tier move with the digest computed by the fused transfer kernel (SURVEY K9 "tier-spill"): copies
  // `len` bytes between this backend (at my_off) and `peer` (at peer_off) in ONE launch and returns the
  // digest of the bytes moved.  NOT_IMPLEMENTED unless this backend owns a device engine and the
  // peer is cuda_accessible(); callers then fall back to read() + write() + CPU checksum.
  virtual ErrorCode device_copy(StorageBackend& peer, bool to_peer, uint64_t my_off, uint64_t peer_off, uint64_t len, ChecksumAlgo algo,
                                uint64_t* digest) {
    (void)peer, (void)to_peer, (void)my_off, (void)peer_off, (void)len, (void)algo, (void)digest;
    return ErrorCode::NOT_IMPLEMENTED;
  }
  // Copies `len` bytes from a peer worker's device slab (identified by its registration key = CUDA IPC handle)
  // into this backend at my_off with the fused kernel; returns the digest of the bytes moved.
  virtual ErrorCode pull_from_peer(const std::vector<uint8_t>& peer_key, uint64_t peer_off, uint64_t my_off, uint64_t len, ChecksumAlgo algo,
                                   uint64_t* digest) {
    (void)peer_key, (void)peer_off, (void)my_off, (void)len, (void)algo, (void)digest;
    return ErrorCode::NOT_IMPLEMENTED;
  }
  uint64_t device_copies() const { return device_copies_; }
  // Accounting for readers that take the bytes through direct_ptr() (zero-copy data server reads).
  void note_read(uint64_t len) { bytes_read_ += len; }
  // Registration key advertised in the pool record ("ucx_rkey_hex"): 8 hex chars of the rkey by
  // default; the GPU tier returns its CUDA IPC handle.
  virtual std::string registration_key_hex() const;

  void set_pool_id(const MemoryPoolId& id) { pool_id_ = id; }
  const MemoryPoolId& pool_id() const { return pool_id_; }
  // Test hook: shortens / moves reservation expiry.
  void set_reservation_ttl_ms(uint64_t ms) { opts_.reservation_ttl_ms = ms; }

 protected:
  StorageBackend(StorageClass sc, uint64_t capacity, BackendOptions opts);
  void init_allocator();  // call from initialize() once capacity is final
  ErrorCode check_range(uint64_t offset, uint64_t len) const;
  // Encryption at rest (BackendOptions::at_rest_key): the cipher of this pool, keyed on first use (the pool id is part of
  // the derivation and is assigned after construction); nullptr = the pool stores plain bytes.  at_rest_check() is for
  // initialize(): a key without a usable libcrypto must fail loudly, not store plain text.
  const net::OffsetCipher* at_rest();
  ErrorCode at_rest_check() const;
  std::once_flag at_rest_once_;
  net::OffsetCipher at_rest_cipher_;

  StorageClass class_;
  uint64_t capacity_;
  BackendOptions opts_;
  MemoryPoolId pool_id_;
  bool initialized_ = false;

  struct Reservation {
    ReservationToken token;
    alloc::Range range;
    std::string owner;  // object key (reserve_shard_at)
  };
  mutable std::mutex mu_;  // guards reservations_/committed_; never held while calling virtuals that lock
  std::unique_ptr<alloc::PoolAllocator> allocator_;
  std::unordered_map<std::string, Reservation> reservations_;
  std::unordered_map<uint64_t, alloc::Range> committed_;  // offset -> extent
  uint64_t next_token_ = 1;
  uint64_t usable_ = 0;  // capacity rounded down to the extent alignment
  std::atomic<uint64_t> device_copies_{0};
  std::atomic<uint64_t> bytes_written_{0}, bytes_read_{0}, io_errors_{0};
};

class RamBackend : public StorageBackend {
 public:
  RamBackend(StorageClass sc, uint64_t capacity, BackendOptions opts = {});
  ~RamBackend() override;
  StorageClass get_storage_class() const override { return class_; }
  uint64_t get_total_capacity() const override { return capacity_; }
  uint64_t get_base_address() const override { return reinterpret_cast<uint64_t>(base_); }
  uint64_t get_rkey() const override { return rkey_; }
  ErrorCode initialize() override;
  void shutdown() override;
  ErrorCode write(uint64_t offset, const void* data, uint64_t len) override;
  ErrorCode read(uint64_t offset, void* data, uint64_t len) override;
  void* direct_ptr(uint64_t offset) override { return base_ ? base_ + offset : nullptr; }
  bool pinned() const { return pinned_; }
  bool cuda_accessible() const override { return pinned_; }
  // "file:<path>" (hex) for shared pools, the 8-hex-digit rkey otherwise.
  std::string registration_key_hex() const override;
  const std::string& shared_path() const { return shared_path_; }

 private:
  uint8_t* base_ = nullptr;
  uint64_t rkey_ = 0;
  bool pinned_ = false;
  int shared_fd_ = -1;
  std::string shared_path_;  // /proc/<pid>/fd/<n> of the backing memfd (shared pools)
};

// Maps a shared RAM pool advertised by another worker ("file:<path>" registration key, raw bytes) into this process.
// Returns nullptr when the key is not a shared-pool key or the object is not reachable from this host / namespace.
void* map_shared_pool(const std::vector<uint8_t>& registration_key, uint64_t size);
void unmap_shared_pool(void* base, uint64_t size);

class MmapDiskBackend : public StorageBackend {
 public:
  MmapDiskBackend(StorageClass sc, uint64_t capacity, BackendOptions opts);
  ~MmapDiskBackend() override;
  StorageClass get_storage_class() const override { return class_; }
  uint64_t get_total_capacity() const override { return capacity_; }
  uint64_t get_base_address() const override { return reinterpret_cast<uint64_t>(map_); }
  uint64_t get_rkey() const override { return rkey_; }
  ErrorCode initialize() override;
  void shutdown() override;
  ErrorCode write(uint64_t offset, const void* data, uint64_t len) override;
  ErrorCode read(uint64_t offset, void* data, uint64_t len) override;
  // (an encrypted pool has no plain bytes anybody could point at)
  void* direct_ptr(uint64_t offset) override { return map_ && opts_.at_rest_key.empty() ? map_ + offset : nullptr; }
  ErrorCode flush() override;
  const std::string& file_path() const { return file_path_; }

 private:
  std::string file_path_;
  int fd_ = -1;
  uint8_t* map_ = nullptr;
  uint64_t rkey_ = 0;
};

// Raw-syscall io_uring wrapper (setup / mmap rings / submit / wait).
class IoUring {
 public:
  IoUring() = default;
  ~IoUring();
  ErrorCode init(uint32_t entries);
  void close();
  bool ok() const { return ring_fd_ >= 0; }
  // Synchronous helpers built on asynchronous submission: queues up to `n` ops, waits for all.
  struct Op {
    bool write;
    int fd;
    void* buf;
    uint32_t len;
    uint64_t offset;
    int32_t result;
  };
  ErrorCode submit_and_wait(std::vector<Op>& ops);
  // Registers one buffer (IORING_REGISTER_BUFFERS): ops whose memory lies inside it are issued as
  // READ_FIXED / WRITE_FIXED, which skips the per-I/O page pinning.
  ErrorCode register_buffer(void* base, size_t len);
  bool has_fixed_buffer() const { return fixed_base_ != nullptr; }
  uint64_t sqes_submitted() const { return submitted_; }
  uint64_t fixed_sqes() const { return fixed_sqes_; }
  static bool supported();

 private:
  int ring_fd_ = -1;
  void* sq_ptr_ = nullptr;
  void* cq_ptr_ = nullptr;
  void* sqes_ = nullptr;
  size_t sq_len_ = 0, cq_len_ = 0, sqes_len_ = 0;
  uint32_t* sq_head_ = nullptr, *sq_tail_ = nullptr, *sq_mask_ = nullptr, *sq_array_ = nullptr;
  uint32_t* cq_head_ = nullptr, *cq_tail_ = nullptr, *cq_mask_ = nullptr;
  void* cqes_ = nullptr;
  uint32_t entries_ = 0;
  uint64_t submitted_ = 0;
  uint64_t fixed_sqes_ = 0;
  uint8_t* fixed_base_ = nullptr;
  size_t fixed_len_ = 0;
  std::mutex mu_;
};

class IoUringDiskBackend : public StorageBackend {
 public:
  IoUringDiskBackend(StorageClass sc, uint64_t capacity, BackendOptions opts);
  ~IoUringDiskBackend() override;
  StorageClass get_storage_class() const override { return class_; }
  uint64_t get_total_capacity() const override { return capacity_; }
  uint64_t get_base_address() const override { return base_tag_; }
  uint64_t get_rkey() const override { return rkey_; }
  ErrorCode initialize() override;
  void shutdown() override;
  ErrorCode commit_shard(const ReservationToken& token) override;
  ErrorCode free_shard(uint64_t remote_addr, uint64_t size) override;
  ErrorCode write(uint64_t offset, const void* data, uint64_t len) override;
  ErrorCode read(uint64_t offset, void* data, uint64_t len) override;
  ErrorCode flush() override;
  const std::string& file_path() const { return file_path_; }
  uint64_t sqes_submitted() const { return ring_.sqes_submitted(); }
  uint64_t fixed_sqes() const { return ring_.fixed_sqes(); }  // READ_FIXED / WRITE_FIXED on the registered staging buffer
  bool using_uring() const { return ring_.ok(); }
  bool using_direct_io() const { return direct_; }
  // Extents recorded in the manifest of a previous run (offset, size, crc32c).
  struct ManifestEntry {
    uint64_t offset, size;
    uint32_t crc;
  };
  std::vector<ManifestEntry> recovered_extents() const { return recovered_; }

 private:
  ErrorCode io(bool is_write, uint64_t offset, void* data, uint64_t len);
  void append_manifest(char op, uint64_t offset, uint64_t size, uint32_t crc);
  void load_manifest();
  std::string dir_, file_path_, manifest_path_;
  int fd_ = -1;
  int manifest_fd_ = -1;
  int buffered_fd_ = -1;  // same file without O_DIRECT: odd offsets and sub-block tails
  bool direct_ = false;
  IoUring ring_;
  uint8_t* staging_ = nullptr;  // aligned bounce buffer for O_DIRECT
  uint64_t staging_bytes_ = 0;
  std::mutex io_mu_;
  uint64_t base_tag_ = 0, rkey_ = 0;
  std::vector<ManifestEntry> recovered_;
};

class CxlMemoryBackend : public StorageBackend {
 public:
  CxlMemoryBackend(StorageClass sc, uint64_t capacity, BackendOptions opts);
  ~CxlMemoryBackend() override;
  StorageClass get_storage_class() const override { return class_; }
  uint64_t get_total_capacity() const override { return capacity_; }
  uint64_t get_base_address() const override { return reinterpret_cast<uint64_t>(base_); }
  uint64_t get_rkey() const override { return rkey_; }
  ErrorCode initialize() override;
  void shutdown() override;
  Result<ReservationToken> reserve_shard(uint64_t size, const std::string& hint = "") override;
  ErrorCode write(uint64_t offset, const void* data, uint64_t len) override;
  ErrorCode read(uint64_t offset, void* data, uint64_t len) override;
  void* direct_ptr(uint64_t offset) override { return base_ ? base_ + offset : nullptr; }
  ErrorCode flush() override;  // persistent mode: msync the DAX mapping (no-op for volatile CXL memory)
  bool is_dax() const { return dax_; }
  bool numa_bound() const { return numa_bound_; }
  uint64_t region_id(uint64_t offset) const { return offset / (opts_.interleave_granularity ? opts_.interleave_granularity : 256); }
  static constexpr uint64_t kCacheLine = 64;

 private:
  uint8_t* base_ = nullptr;
  uint64_t map_len_ = 0;
  int fd_ = -1;
  bool dax_ = false;
  bool numa_bound_ = false;
  uint64_t rkey_ = 0;
};

// Hook through which the CUDA side registers the GPU tier (keeps this library CUDA-free).
// Page-locking hooks installed by the CUDA side (cudaHostRegister / cudaHostUnregister): a DRAM pool created
// with BackendOptions::pin_memory becomes addressable by the fused kernels of this process.
struct HostPinHooks {
  std::function<bool(void*, uint64_t)> pin;
  std::function<void(void*)> unpin;
};
void set_host_pin_hooks(HostPinHooks h);
HostPinHooks host_pin_hooks();

// Process-local registry of pinned (CUDA-registered) RAM pools: a GPU client living in the same process addresses
// them directly instead of mapping the backing memfd a second time.
struct LocalHostPool {
  void* base = nullptr;
  uint64_t size = 0;
};
void register_local_host_pool(const std::string& pool_id, void* base, uint64_t size);
void unregister_local_host_pool(const std::string& pool_id);
bool find_local_host_pool(const std::string& pool_id, LocalHostPool* out);

using GpuBackendFactory = std::function<std::unique_ptr<StorageBackend>(uint64_t capacity, const BackendOptions&)>;
void set_gpu_backend_factory(GpuBackendFactory f);

// nullptr when the class cannot be built on this host (e.g. RAM_GPU without a GPU).
std::unique_ptr<StorageBackend> create_storage_backend(StorageClass sc, uint64_t capacity, const BackendOptions& opts = {});

}  // namespace bb::worker
