#include "worker/storage_backend.h"

#include <fcntl.h>
#include <linux/io_uring.h>
#include <sys/mman.h>
#include <sys/uio.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <sys/types.h>
#include <unistd.h>

#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <filesystem>

#include "common/checksum.h"
#include "common/log.h"
#include "common/sha256.h"

namespace bb::worker {

namespace {
GpuBackendFactory g_gpu_factory;
std::mutex g_factory_mu;
uint64_t fnv64(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
  return h;
}
constexpr uint64_t kBlock = 4096;
}  // namespace

// ================================================================ base
StorageBackend::StorageBackend(StorageClass sc, uint64_t capacity, BackendOptions opts)
    : class_(sc), capacity_(capacity), opts_(std::move(opts)) {}

void StorageBackend::init_allocator() {
  MemoryPool p;
  p.id = pool_id_.empty() ? "backend" : pool_id_;
  p.size = capacity_;
  p.storage_class = class_;
  const uint64_t align = is_disk_class(class_) ? kBlock : (class_ == StorageClass::CXL_MEMORY || class_ == StorageClass::CXL_TYPE2_DEVICE) ? 64 : 256;
  allocator_ = std::make_unique<alloc::PoolAllocator>(p, align);
  usable_ = allocator_->total_free();
}

const net::OffsetCipher* StorageBackend::at_rest() {
  if (opts_.at_rest_key.empty()) return nullptr;
  std::call_once(at_rest_once_, [this] {
    // one key and one counter-block prefix per (passphrase, pool): the same passphrase on two pools never reuses a key stream
    const std::string& scope = opts_.at_rest_scope.empty() ? pool_id_ : opts_.at_rest_scope;
    const Sha256Digest k = hmac_sha256(opts_.at_rest_key, "bb-at-rest-key:" + scope);
    const Sha256Digest n = hmac_sha256(opts_.at_rest_key, "bb-at-rest-nonce:" + scope);
    if (!at_rest_cipher_.set_key(k.data(), n.data()))
      BB_LOG(ERROR) << "pool " << pool_id_ << ": at_rest_key is set but libcrypto's AES-CTR is not available; refusing I/O";
  });
  return &at_rest_cipher_;  // not ready() -> crypt() fails -> I/O fails: never falls back to plain text
}

ErrorCode StorageBackend::at_rest_check() const {
  if (opts_.at_rest_key.empty()) return ErrorCode::OK;
  std::string why;
  if (!net::Aead::available(&why)) {
    BB_LOG(ERROR) << "encryption at rest requested but unavailable: " << why;
    return ErrorCode::CONFIG_ERROR;
  }
  return ErrorCode::OK;
}

ErrorCode StorageBackend::check_range(uint64_t offset, uint64_t len) const {
  if (!initialized_) return ErrorCode::INVALID_STATE;
  if (offset > capacity_ || len > capacity_ - offset) return ErrorCode::MEMORY_ACCESS_ERROR;
  return ErrorCode::OK;
}

uint64_t StorageBackend::get_used_capacity() const {
  if (!allocator_) return 0;
  const uint64_t usable = usable_;  // extents are alignment-granular
  const uint64_t free_b = allocator_->total_free();
  return usable > free_b ? usable - free_b : 0;
}
uint64_t StorageBackend::get_available_capacity() const { return allocator_ ? allocator_->total_free() : 0; }

Result<ReservationToken> StorageBackend::reserve_shard(uint64_t size, const std::string& hint) {
  if (!initialized_) return ErrorCode::INVALID_STATE;
  if (size == 0) return ErrorCode::INVALID_PARAMETERS;
  // reclaim expired reservations first so that abandoned puts cannot exhaust the tier
  {
    std::vector<Reservation> expired;
    const auto now = std::chrono::system_clock::now();
    std::lock_guard<std::mutex> lk(mu_);
    for (auto it = reservations_.begin(); it != reservations_.end();) {
      if (it->second.token.expires_at <= now) {
        expired.push_back(it->second);
        it = reservations_.erase(it);
      } else {
        ++it;
      }
    }
    for (const auto& r : expired) allocator_->free(r.range);
  }
  auto range = allocator_->allocate(size, true);
  if (!range) return ErrorCode::OUT_OF_MEMORY;
  Reservation r;
  r.range = *range;
  r.token.pool_id = pool_id_;
  r.token.remote_addr = get_base_address() + range->offset;
  r.token.rkey = get_rkey();
  r.token.size = size;
  r.token.expires_at = std::chrono::system_clock::now() + std::chrono::milliseconds(opts_.reservation_ttl_ms);
  std::lock_guard<std::mutex> lk(mu_);
  r.token.token_id = (hint.empty() ? std::string("tok") : hint) + "-" + std::to_string(next_token_++);
  reservations_[r.token.token_id] = r;
  return r.token;
}

Result<ReservationToken> StorageBackend::reserve_shard_at(uint64_t offset, uint64_t size, const std::string& owner, uint64_t ttl_ms) {
  if (!initialized_) return ErrorCode::INVALID_STATE;
  if (size == 0) return ErrorCode::INVALID_PARAMETERS;
  BB_TRY(check_range(offset, size));
  if (!allocator_->allocate_at(offset, size)) return ErrorCode::ALLOCATION_FAILED;
  Reservation r;
  r.range = alloc::Range(offset, allocator_->aligned(size));
  r.owner = owner;
  r.token.pool_id = pool_id_;
  r.token.remote_addr = get_base_address() + offset;
  r.token.rkey = get_rkey();
  r.token.size = size;
  r.token.expires_at = std::chrono::system_clock::now() + std::chrono::milliseconds(ttl_ms ? ttl_ms : opts_.reservation_ttl_ms);
  std::lock_guard<std::mutex> lk(mu_);
  r.token.token_id = pool_id_ + "#" + std::to_string(next_token_++);
  reservations_[r.token.token_id] = r;
  return r.token;
}

ErrorCode StorageBackend::commit_shard_id(const std::string& token_id) {
  ReservationToken t;
  t.token_id = token_id;
  return commit_shard(t);
}
ErrorCode StorageBackend::abort_shard_id(const std::string& token_id) {
  ReservationToken t;
  t.token_id = token_id;
  return abort_shard(t);
}

std::vector<std::pair<std::string, std::string>> StorageBackend::reap_expired_reservations() {
  std::vector<std::pair<std::string, std::string>> out;
  if (!initialized_) return out;
  const auto now = std::chrono::system_clock::now();
  std::lock_guard<std::mutex> lk(mu_);
  for (auto it = reservations_.begin(); it != reservations_.end();) {
    if (it->second.token.expires_at <= now) {
      allocator_->free(it->second.range);
      out.emplace_back(it->first, it->second.owner);
      it = reservations_.erase(it);
    } else {
      ++it;
    }
  }
  return out;
}

ErrorCode StorageBackend::commit_shard(const ReservationToken& token) {
  Reservation r;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = reservations_.find(token.token_id);
    if (it == reservations_.end()) return ErrorCode::INVALID_PARAMETERS;
    r = it->second;
    reservations_.erase(it);
    if (r.token.expires_at <= std::chrono::system_clock::now()) {
      allocator_->free(r.range);
      return ErrorCode::OPERATION_TIMEOUT;
    }
    committed_[r.range.offset] = r.range;
  }
  return ErrorCode::OK;
}

ErrorCode StorageBackend::abort_shard(const ReservationToken& token) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = reservations_.find(token.token_id);
  if (it == reservations_.end()) return ErrorCode::INVALID_PARAMETERS;
  allocator_->free(it->second.range);
  reservations_.erase(it);
  return ErrorCode::OK;
}

ErrorCode StorageBackend::free_shard(uint64_t remote_addr, uint64_t size) {
  const uint64_t base = get_base_address();
  if (remote_addr < base) return ErrorCode::MEMORY_ACCESS_ERROR;
  const uint64_t off = remote_addr - base;
  std::lock_guard<std::mutex> lk(mu_);
  auto it = committed_.find(off);
  if (it == committed_.end()) return ErrorCode::OBJECT_NOT_FOUND;
  if (allocator_->aligned(size) != it->second.length) return ErrorCode::INVALID_PARAMETERS;
  allocator_->free(it->second);
  committed_.erase(it);
  return ErrorCode::OK;
}

std::string StorageBackend::registration_key_hex() const {
  char buf[20];
  std::snprintf(buf, sizeof buf, "%08llx", static_cast<unsigned long long>(get_rkey() & 0xFFFFFFFFull));
  return buf;
}

StorageStats StorageBackend::get_stats() const {
  StorageStats s;
  s.total_capacity = capacity_;
  if (allocator_) {
    s.available_capacity = allocator_->total_free();
    s.used_capacity = usable_ > s.available_capacity ? usable_ - s.available_capacity : 0;
    s.fragmentation = allocator_->fragmentation_ratio();
  }
  {
    std::lock_guard<std::mutex> lk(mu_);
    s.num_reservations = reservations_.size();
    s.num_committed_shards = committed_.size();
  }
  s.utilization = capacity_ ? static_cast<double>(s.used_capacity) / static_cast<double>(capacity_) : 0.0;
  s.bytes_written = bytes_written_.load();
  s.bytes_read = bytes_read_.load();
  s.io_errors = io_errors_.load();
  return s;
}

// ================================================================ RamBackend
RamBackend::RamBackend(StorageClass sc, uint64_t capacity, BackendOptions opts) : StorageBackend(sc, capacity, std::move(opts)) {}
RamBackend::~RamBackend() { shutdown(); }

ErrorCode RamBackend::initialize() {
  if (initialized_) return ErrorCode::OK;
  if (capacity_ == 0) return ErrorCode::INVALID_PARAMETERS;
  void* p = MAP_FAILED;
  if (opts_.shared_memory) {
    // memfd: lives as long as this process, needs no name in /dev/shm (whose mount is tiny in containers) and is
    // reachable by same-user peers through /proc/<pid>/fd/<n>.
    shared_fd_ = static_cast<int>(::syscall(SYS_memfd_create, "bb-ram-pool", 1u /* MFD_CLOEXEC */));
    if (shared_fd_ >= 0 && ::ftruncate(shared_fd_, static_cast<off_t>(capacity_)) == 0)
      p = ::mmap(nullptr, capacity_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_NORESERVE, shared_fd_, 0);
    if (p == MAP_FAILED) {
      BB_LOG(WARNING) << "RAM pool " << pool_id_ << ": shared_memory requested but memfd setup failed (" << std::strerror(errno)
                      << "); falling back to a private mapping";
      if (shared_fd_ >= 0) ::close(shared_fd_);
      shared_fd_ = -1;
    } else {
      shared_path_ = "/proc/" + std::to_string(::getpid()) + "/fd/" + std::to_string(shared_fd_);
    }
  }
  if (p == MAP_FAILED) p = ::mmap(nullptr, capacity_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return ErrorCode::OUT_OF_MEMORY;
  base_ = static_cast<uint8_t*>(p);
  rkey_ = fnv64(pool_id_) & 0xFFFFFFFFull;
  if (opts_.pin_memory) {
    HostPinHooks h = host_pin_hooks();
    if (h.pin && h.pin(base_, capacity_)) {
      pinned_ = true;
      register_local_host_pool(pool_id_, base_, capacity_);
    } else BB_LOG(INFO) << "RAM pool " << pool_id_ << ": pin_memory requested but no CUDA pin hook (or registration failed); tier moves use the staged path";
  }
  init_allocator();
  initialized_ = true;
  return ErrorCode::OK;
}

void RamBackend::shutdown() {
  if (base_ && pinned_) {
    HostPinHooks h = host_pin_hooks();
    unregister_local_host_pool(pool_id_);
    if (h.unpin) h.unpin(base_);
    pinned_ = false;
  }
  if (base_) ::munmap(base_, capacity_);
  base_ = nullptr;
  if (shared_fd_ >= 0) ::close(shared_fd_);
  shared_fd_ = -1;
  shared_path_.clear();
  initialized_ = false;
}

std::string RamBackend::registration_key_hex() const {
  if (shared_path_.empty()) return StorageBackend::registration_key_hex();
  const std::string key = "file:" + shared_path_;
  static const char* hexd = "0123456789abcdef";
  std::string out;
  out.reserve(key.size() * 2);
  for (unsigned char c : key) {
    out.push_back(hexd[c >> 4]);
    out.push_back(hexd[c & 15]);
  }
  return out;
}

namespace {
std::mutex g_host_pools_mu;
std::unordered_map<std::string, LocalHostPool> g_host_pools;
}  // namespace
void register_local_host_pool(const std::string& pool_id, void* base, uint64_t size) {
  std::lock_guard<std::mutex> lk(g_host_pools_mu);
  g_host_pools[pool_id] = LocalHostPool{base, size};
}
void unregister_local_host_pool(const std::string& pool_id) {
  std::lock_guard<std::mutex> lk(g_host_pools_mu);
  g_host_pools.erase(pool_id);
}
bool find_local_host_pool(const std::string& pool_id, LocalHostPool* out) {
  std::lock_guard<std::mutex> lk(g_host_pools_mu);
  auto it = g_host_pools.find(pool_id);
  if (it == g_host_pools.end()) return false;
  if (out) *out = it->second;
  return true;
}

void* map_shared_pool(const std::vector<uint8_t>& key, uint64_t size) {
  static const char kPrefix[] = "file:";
  if (size == 0 || key.size() <= 5 || std::memcmp(key.data(), kPrefix, 5) != 0) return nullptr;
  const std::string path(reinterpret_cast<const char*>(key.data()) + 5, key.size() - 5);
  if (path.find('\0') != std::string::npos) return nullptr;
  const int fd = ::open(path.c_str(), O_RDWR | O_CLOEXEC);
  if (fd < 0) return nullptr;  // another host / pid namespace / user: the data server path is used instead
  struct stat st {};
  void* p = MAP_FAILED;
  if (::fstat(fd, &st) == 0 && static_cast<uint64_t>(st.st_size) >= size) p = ::mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  ::close(fd);
  return p == MAP_FAILED ? nullptr : p;
}

void unmap_shared_pool(void* base, uint64_t size) {
  if (base) ::munmap(base, size);
}

ErrorCode RamBackend::write(uint64_t offset, const void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  std::memcpy(base_ + offset, data, len);
  bytes_written_ += len;
  return ErrorCode::OK;
}
ErrorCode RamBackend::read(uint64_t offset, void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  std::memcpy(data, base_ + offset, len);
  bytes_read_ += len;
  return ErrorCode::OK;
}

// ================================================================ MmapDiskBackend
MmapDiskBackend::MmapDiskBackend(StorageClass sc, uint64_t capacity, BackendOptions opts) : StorageBackend(sc, capacity, std::move(opts)) {}
MmapDiskBackend::~MmapDiskBackend() { shutdown(); }

ErrorCode MmapDiskBackend::initialize() {
  if (initialized_) return ErrorCode::OK;
  if (!is_disk_class(class_)) return ErrorCode::INVALID_PARAMETERS;
  BB_TRY(at_rest_check());
  std::error_code fe;
  const std::string dir = (opts_.mount_path.empty() ? std::string("/tmp") : opts_.mount_path) + "/blackbird_mmap_storage";
  std::filesystem::create_directories(dir, fe);
  if (fe) return ErrorCode::IO_ERROR;
  file_path_ = dir + "/" + (pool_id_.empty() ? std::string("mmap_storage") : pool_id_) + ".dat";
  fd_ = ::open(file_path_.c_str(), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
  if (fd_ < 0) return ErrorCode::IO_ERROR;
  if (::ftruncate(fd_, static_cast<off_t>(capacity_)) != 0) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::INSUFFICIENT_SPACE;
  }
  void* p = ::mmap(nullptr, capacity_, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
  if (p == MAP_FAILED) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::OUT_OF_MEMORY;
  }
  map_ = static_cast<uint8_t*>(p);
  ::madvise(map_, capacity_, MADV_RANDOM);
  rkey_ = fnv64(file_path_) & 0xFFFFFFFFull;
  init_allocator();
  initialized_ = true;
  return ErrorCode::OK;
}

void MmapDiskBackend::shutdown() {
  if (map_) {
    ::msync(map_, capacity_, MS_ASYNC);
    ::munmap(map_, capacity_);
  }
  if (fd_ >= 0) ::close(fd_);
  map_ = nullptr;
  fd_ = -1;
  initialized_ = false;
}

ErrorCode MmapDiskBackend::write(uint64_t offset, const void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  if (const net::OffsetCipher* c = at_rest()) {
    if (!c->crypt(offset, data, map_ + offset, len)) return ErrorCode::IO_ERROR;  // encrypts straight into the mapping
  } else {
    std::memcpy(map_ + offset, data, len);
  }
  bytes_written_ += len;
  return ErrorCode::OK;
}
ErrorCode MmapDiskBackend::read(uint64_t offset, void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  if (const net::OffsetCipher* c = at_rest()) {
    if (!c->crypt(offset, map_ + offset, data, len)) return ErrorCode::IO_ERROR;
  } else {
    std::memcpy(data, map_ + offset, len);
  }
  bytes_read_ += len;
  return ErrorCode::OK;
}
ErrorCode MmapDiskBackend::flush() {
  if (!map_) return ErrorCode::INVALID_STATE;
  return ::msync(map_, capacity_, MS_SYNC) == 0 ? ErrorCode::OK : ErrorCode::IO_ERROR;
}

// ================================================================ IoUring (raw syscalls)
IoUring::~IoUring() { close(); }

bool IoUring::supported() {
  IoUring r;
  return r.init(2) == ErrorCode::OK;
}

ErrorCode IoUring::init(uint32_t entries) {
  close();
  io_uring_params p;
  std::memset(&p, 0, sizeof p);
  int fd = static_cast<int>(::syscall(__NR_io_uring_setup, entries, &p));
  if (fd < 0) return ErrorCode::NOT_IMPLEMENTED;
  ring_fd_ = fd;
  entries_ = p.sq_entries;
  sq_len_ = p.sq_off.array + p.sq_entries * sizeof(uint32_t);
  cq_len_ = p.cq_off.cqes + p.cq_entries * sizeof(io_uring_cqe);
  const bool single = p.features & IORING_FEAT_SINGLE_MMAP;
  if (single) sq_len_ = cq_len_ = std::max(sq_len_, cq_len_);
  sq_ptr_ = ::mmap(nullptr, sq_len_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQ_RING);
  if (sq_ptr_ == MAP_FAILED) {
    sq_ptr_ = nullptr;
    close();
    return ErrorCode::OUT_OF_MEMORY;
  }
  if (single) {
    cq_ptr_ = sq_ptr_;
  } else {
    cq_ptr_ = ::mmap(nullptr, cq_len_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_CQ_RING);
    if (cq_ptr_ == MAP_FAILED) {
      cq_ptr_ = nullptr;
      close();
      return ErrorCode::OUT_OF_MEMORY;
    }
  }
  sqes_len_ = p.sq_entries * sizeof(io_uring_sqe);
  sqes_ = ::mmap(nullptr, sqes_len_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_POPULATE, fd, IORING_OFF_SQES);
  if (sqes_ == MAP_FAILED) {
    sqes_ = nullptr;
    close();
    return ErrorCode::OUT_OF_MEMORY;
  }
  auto* sq = static_cast<uint8_t*>(sq_ptr_);
  auto* cq = static_cast<uint8_t*>(cq_ptr_);
  sq_head_ = reinterpret_cast<uint32_t*>(sq + p.sq_off.head);
  sq_tail_ = reinterpret_cast<uint32_t*>(sq + p.sq_off.tail);
  sq_mask_ = reinterpret_cast<uint32_t*>(sq + p.sq_off.ring_mask);
  sq_array_ = reinterpret_cast<uint32_t*>(sq + p.sq_off.array);
  cq_head_ = reinterpret_cast<uint32_t*>(cq + p.cq_off.head);
  cq_tail_ = reinterpret_cast<uint32_t*>(cq + p.cq_off.tail);
  cq_mask_ = reinterpret_cast<uint32_t*>(cq + p.cq_off.ring_mask);
  cqes_ = cq + p.cq_off.cqes;
  return ErrorCode::OK;
}

void IoUring::close() {
  if (sqes_) ::munmap(sqes_, sqes_len_);
  if (cq_ptr_ && cq_ptr_ != sq_ptr_) ::munmap(cq_ptr_, cq_len_);
  if (sq_ptr_) ::munmap(sq_ptr_, sq_len_);
  if (ring_fd_ >= 0) ::close(ring_fd_);
  sqes_ = cq_ptr_ = sq_ptr_ = nullptr;
  ring_fd_ = -1;
  fixed_base_ = nullptr;
  fixed_len_ = 0;
}

ErrorCode IoUring::register_buffer(void* base, size_t len) {
  std::lock_guard<std::mutex> lk(mu_);
  if (ring_fd_ < 0) return ErrorCode::INVALID_STATE;
  struct iovec iov;
  iov.iov_base = base;
  iov.iov_len = len;
  long rc = ::syscall(__NR_io_uring_register, ring_fd_, IORING_REGISTER_BUFFERS, &iov, 1);
  if (rc < 0) return ErrorCode::NOT_IMPLEMENTED;  // e.g. RLIMIT_MEMLOCK: plain READ/WRITE still work
  fixed_base_ = static_cast<uint8_t*>(base);
  fixed_len_ = len;
  return ErrorCode::OK;
}

ErrorCode IoUring::submit_and_wait(std::vector<Op>& ops) {
  std::lock_guard<std::mutex> lk(mu_);
  if (ring_fd_ < 0) return ErrorCode::INVALID_STATE;
  size_t done = 0;
  while (done < ops.size()) {
    const uint32_t n = static_cast<uint32_t>(std::min<size_t>(entries_, ops.size() - done));
    uint32_t tail = *sq_tail_;
    auto* sqes = static_cast<io_uring_sqe*>(sqes_);
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t idx = tail & *sq_mask_;
      io_uring_sqe& e = sqes[idx];
      std::memset(&e, 0, sizeof e);
      Op& op = ops[done + i];
      auto* b8 = static_cast<uint8_t*>(op.buf);
      const bool fixed = fixed_base_ && b8 >= fixed_base_ && b8 + op.len <= fixed_base_ + fixed_len_;
      if (fixed) {
        e.opcode = op.write ? IORING_OP_WRITE_FIXED : IORING_OP_READ_FIXED;
        e.buf_index = 0;
        ++fixed_sqes_;
      } else {
        e.opcode = op.write ? IORING_OP_WRITE : IORING_OP_READ;
      }
      e.fd = op.fd;
      e.addr = reinterpret_cast<uint64_t>(op.buf);
      e.len = op.len;
      e.off = op.offset;
      e.user_data = done + i;
      sq_array_[idx] = idx;
      ++tail;
    }
    __atomic_store_n(sq_tail_, tail, __ATOMIC_RELEASE);
    uint32_t reaped = 0;
    uint32_t to_submit = n;
    while (reaped < n) {
      long rc = ::syscall(__NR_io_uring_enter, ring_fd_, to_submit, n - reaped, IORING_ENTER_GETEVENTS, nullptr, 0);
      if (rc < 0) {
        if (errno == EINTR) continue;
        return ErrorCode::IO_ERROR;
      }
      submitted_ += to_submit ? static_cast<uint64_t>(rc) : 0;
      to_submit = 0;
      uint32_t head = *cq_head_;
      const uint32_t ctail = __atomic_load_n(cq_tail_, __ATOMIC_ACQUIRE);
      auto* cqes = static_cast<io_uring_cqe*>(cqes_);
      while (head != ctail) {
        const io_uring_cqe& c = cqes[head & *cq_mask_];
        if (c.user_data < ops.size()) ops[c.user_data].result = c.res;
        ++head;
        ++reaped;
      }
      __atomic_store_n(cq_head_, head, __ATOMIC_RELEASE);
    }
    done += n;
  }
  return ErrorCode::OK;
}

// ================================================================ IoUringDiskBackend
IoUringDiskBackend::IoUringDiskBackend(StorageClass sc, uint64_t capacity, BackendOptions opts)
    : StorageBackend(sc, capacity, std::move(opts)) {}
IoUringDiskBackend::~IoUringDiskBackend() { shutdown(); }

ErrorCode IoUringDiskBackend::initialize() {
  if (initialized_) return ErrorCode::OK;
  if (!is_disk_class(class_)) return ErrorCode::INVALID_PARAMETERS;
  BB_TRY(at_rest_check());
  std::error_code fe;
  dir_ = (opts_.mount_path.empty() ? std::string("/tmp") : opts_.mount_path) + "/blackbird_storage";
  std::filesystem::create_directories(dir_, fe);
  if (fe) return ErrorCode::IO_ERROR;
  const std::string stem = pool_id_.empty() ? std::string("pool") : pool_id_;
  file_path_ = dir_ + "/" + stem + ".bbdata";
  manifest_path_ = dir_ + "/" + stem + ".manifest";
  // O_DIRECT for flash tiers when the filesystem allows it (tmpfs does not)
  if (class_ == StorageClass::NVME || class_ == StorageClass::SSD) {
    fd_ = ::open(file_path_.c_str(), O_RDWR | O_CREAT | O_CLOEXEC | O_DIRECT, 0644);
    direct_ = fd_ >= 0;
    if (direct_) buffered_fd_ = ::open(file_path_.c_str(), O_RDWR | O_CLOEXEC);
  }
  if (fd_ < 0) fd_ = ::open(file_path_.c_str(), O_RDWR | O_CREAT | O_CLOEXEC, 0644);
  if (fd_ < 0) return ErrorCode::IO_ERROR;
  if (::ftruncate(fd_, static_cast<off_t>(capacity_)) != 0) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::INSUFFICIENT_SPACE;
  }
  load_manifest();
  manifest_fd_ = ::open(manifest_path_.c_str(), O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0644);
  if (ring_.init(std::max<uint32_t>(2, opts_.queue_depth)) != ErrorCode::OK)
    BB_LOG(WARNING) << "io_uring unavailable; " << pool_id_ << " falls back to pread/pwrite";
  staging_bytes_ = std::min<uint64_t>(16ull << 20, std::max<uint64_t>(1ull << 20, static_cast<uint64_t>(opts_.queue_depth) * (256ull << 10)));
  if (::posix_memalign(reinterpret_cast<void**>(&staging_), kBlock, staging_bytes_) != 0) return ErrorCode::OUT_OF_MEMORY;
  if (ring_.ok() && ring_.register_buffer(staging_, staging_bytes_) != ErrorCode::OK)
    BB_LOG(INFO) << pool_id_ << ": staging buffer not registered with io_uring (memlock limit?); using READ/WRITE instead of *_FIXED";
  base_tag_ = (fnv64(file_path_) >> 32) << 32;  // synthetic "address space" of this file
  rkey_ = fnv64(manifest_path_) & 0xFFFFFFFFull;
  init_allocator();
  // extents that survived a restart stay reserved
  for (const auto& e : recovered_)
    if (allocator_->allocate_at(e.offset, e.size)) committed_[e.offset] = alloc::Range(e.offset, allocator_->aligned(e.size));
  initialized_ = true;
  return ErrorCode::OK;
}

void IoUringDiskBackend::shutdown() {
  if (fd_ >= 0) {
    ::fsync(fd_);
    ::close(fd_);
  }
  if (manifest_fd_ >= 0) ::close(manifest_fd_);
  if (buffered_fd_ >= 0) ::close(buffered_fd_);
  buffered_fd_ = -1;
  ring_.close();
  std::free(staging_);
  staging_ = nullptr;
  fd_ = manifest_fd_ = -1;
  initialized_ = false;
}

void IoUringDiskBackend::append_manifest(char op, uint64_t offset, uint64_t size, uint32_t crc) {
  if (manifest_fd_ < 0) return;
  char line[96];
  const int n = std::snprintf(line, sizeof line, "%c %llu %llu %08x\n", op, static_cast<unsigned long long>(offset),
                              static_cast<unsigned long long>(size), crc);
  if (n > 0 && ::write(manifest_fd_, line, static_cast<size_t>(n)) != n) io_errors_++;
}

void IoUringDiskBackend::load_manifest() {
  FILE* f = std::fopen(manifest_path_.c_str(), "r");
  if (!f) return;
  std::unordered_map<uint64_t, ManifestEntry> live;
  char op;
  unsigned long long off, size;
  unsigned crc;
  while (std::fscanf(f, " %c %llu %llu %x", &op, &off, &size, &crc) == 4) {
    if (op == 'C') live[off] = ManifestEntry{off, size, crc};
    else if (op == 'F') live.erase(off);
  }
  std::fclose(f);
  for (const auto& [o, e] : live) recovered_.push_back(e);
}

ErrorCode IoUringDiskBackend::io(bool is_write, uint64_t offset, void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  if (len == 0) return ErrorCode::OK;
  std::lock_guard<std::mutex> lk(io_mu_);
  const bool aligned = (offset % kBlock) == 0;
  auto* user = static_cast<uint8_t*>(data);
  // buffered side door for what O_DIRECT cannot express (odd offsets, sub-block tails); the kernel keeps it coherent with
  // the direct fd (direct I/O writes back / invalidates the page-cache range it touches)
  auto buffered = [&](uint64_t at, uint8_t* p, uint64_t n) {
    const int bfd = buffered_fd_ >= 0 ? buffered_fd_ : fd_;
    for (uint64_t done = 0; done < n;) {
      const ssize_t rc = is_write ? ::pwrite(bfd, p + done, n - done, static_cast<off_t>(at + done))
                                  : ::pread(bfd, p + done, n - done, static_cast<off_t>(at + done));
      if (rc <= 0) return false;
      done += static_cast<uint64_t>(rc);
    }
    return true;
  };
  uint64_t pos = 0;
  while (pos < len) {
    uint64_t chunk = std::min<uint64_t>(len - pos, staging_bytes_);
    if (direct_ && (!aligned || offset + pos + (chunk + kBlock - 1) / kBlock * kBlock > capacity_)) {
      if (!buffered(offset + pos, user + pos, chunk)) {
        io_errors_++;
        return ErrorCode::IO_ERROR;
      }
      pos += chunk;
      continue;
    }
    // A direct READ may run past the request into the bounce buffer; a direct WRITE must not: Keystone packs extents at
    // 256 B granularity, so the bytes after `len` can be a neighbour's.  (A zero-padded tail block used to wipe up to 4 KiB
    // of the object placed behind this one.)  Write whole blocks directly and the sub-block tail through the page cache.
    if (direct_ && is_write && chunk % kBlock) {
      const uint64_t body = chunk / kBlock * kBlock;
      if (body == 0) {
        if (!buffered(offset + pos, user + pos, chunk)) {
          io_errors_++;
          return ErrorCode::IO_ERROR;
        }
        pos += chunk;
        continue;
      }
      chunk = body;  // the tail is picked up by the next turn of the loop
    }
    const uint64_t padded = direct_ ? (chunk + kBlock - 1) / kBlock * kBlock : chunk;
    if (is_write) {
      std::memcpy(staging_, user + pos, chunk);
      if (padded > chunk) std::memset(staging_ + chunk, 0, padded - chunk);
    }
    if (ring_.ok()) {
      // split the chunk into <= queue_depth SQEs of 256 KiB so the device sees parallel I/O
      std::vector<IoUring::Op> ops;
      const uint64_t piece = 256ull << 10;
      for (uint64_t o = 0; o < padded; o += piece) {
        IoUring::Op op;
        op.write = is_write;
        op.fd = fd_;
        op.buf = staging_ + o;
        op.len = static_cast<uint32_t>(std::min<uint64_t>(piece, padded - o));
        op.offset = offset + pos + o;
        op.result = 0;
        ops.push_back(op);
      }
      if (ring_.submit_and_wait(ops) != ErrorCode::OK) {
        io_errors_++;
        return ErrorCode::IO_ERROR;
      }
      for (const auto& op : ops)
        if (op.result < 0 || (is_write && static_cast<uint32_t>(op.result) != op.len)) {
          io_errors_++;
          BB_LOG(ERROR) << "io_uring " << (is_write ? "write" : "read") << " failed: " << std::strerror(-op.result);
          return ErrorCode::IO_ERROR;
        }
    } else {
      ssize_t rc = is_write ? ::pwrite(fd_, staging_, padded, static_cast<off_t>(offset + pos))
                            : ::pread(fd_, staging_, padded, static_cast<off_t>(offset + pos));
      if (rc < static_cast<ssize_t>(is_write ? padded : chunk)) {
        io_errors_++;
        return ErrorCode::IO_ERROR;
      }
    }
    if (!is_write) std::memcpy(user + pos, staging_, chunk);
    pos += chunk;
  }
  (is_write ? bytes_written_ : bytes_read_) += len;
  return ErrorCode::OK;
}

ErrorCode IoUringDiskBackend::write(uint64_t offset, const void* data, uint64_t len) {
  if (const net::OffsetCipher* c = at_rest()) {  // what reaches the ring (and the file) is cipher text
    BB_TRY(check_range(offset, len));
    constexpr uint64_t kPiece = 8ull << 20;
    std::vector<uint8_t> enc(std::min<uint64_t>(len, kPiece));
    for (uint64_t pos = 0; pos < len; pos += kPiece) {
      const uint64_t n = std::min<uint64_t>(kPiece, len - pos);
      if (!c->crypt(offset + pos, static_cast<const uint8_t*>(data) + pos, enc.data(), n)) return ErrorCode::IO_ERROR;
      BB_TRY(io(true, offset + pos, enc.data(), n));
    }
    return ErrorCode::OK;
  }
  return io(true, offset, const_cast<void*>(data), len);
}
ErrorCode IoUringDiskBackend::read(uint64_t offset, void* data, uint64_t len) {
  BB_TRY(io(false, offset, data, len));
  if (const net::OffsetCipher* c = at_rest())
    if (!c->crypt(offset, data, data, len)) return ErrorCode::IO_ERROR;  // in place
  return ErrorCode::OK;
}
ErrorCode IoUringDiskBackend::flush() { return fd_ >= 0 && ::fsync(fd_) == 0 ? ErrorCode::OK : ErrorCode::IO_ERROR; }

ErrorCode IoUringDiskBackend::commit_shard(const ReservationToken& token) {
  ErrorCode ec = StorageBackend::commit_shard(token);
  if (ec != ErrorCode::OK) return ec;
  // record the extent with the CRC of what is on disk so a restarted worker can verify it
  const uint64_t off = token.remote_addr - get_base_address();
  std::vector<uint8_t> buf(std::min<uint64_t>(token.size, 8ull << 20));
  uint32_t crc = 0;
  for (uint64_t p = 0; p < token.size; p += buf.size()) {
    const uint64_t n = std::min<uint64_t>(buf.size(), token.size - p);
    if (read(off + p, buf.data(), n) != ErrorCode::OK) break;
    crc = crc32c(buf.data(), n, crc);
  }
  append_manifest('C', off, token.size, crc);
  return ErrorCode::OK;
}

ErrorCode IoUringDiskBackend::free_shard(uint64_t remote_addr, uint64_t size) {
  ErrorCode ec = StorageBackend::free_shard(remote_addr, size);
  if (ec == ErrorCode::OK) append_manifest('F', remote_addr - get_base_address(), size, 0);
  return ec;
}

// ================================================================ CxlMemoryBackend
CxlMemoryBackend::CxlMemoryBackend(StorageClass sc, uint64_t capacity, BackendOptions opts)
    : StorageBackend(sc, (capacity / kCacheLine) * kCacheLine, std::move(opts)) {}
CxlMemoryBackend::~CxlMemoryBackend() { shutdown(); }

ErrorCode CxlMemoryBackend::initialize() {
  if (initialized_) return ErrorCode::OK;
  if (class_ != StorageClass::CXL_MEMORY && class_ != StorageClass::CXL_TYPE2_DEVICE) return ErrorCode::INVALID_ARGUMENT;
  if (capacity_ == 0) return ErrorCode::INVALID_ARGUMENT;
  void* p = MAP_FAILED;
  if (!opts_.mount_path.empty()) {
    fd_ = ::open(opts_.mount_path.c_str(), O_RDWR | O_CLOEXEC);
    if (fd_ >= 0) {
      p = ::mmap(nullptr, capacity_, PROT_READ | PROT_WRITE, MAP_SHARED, fd_, 0);
      dax_ = p != MAP_FAILED;
      if (!dax_) {
        ::close(fd_);
        fd_ = -1;
      }
    }
  }
  if (p == MAP_FAILED) {
    // no /dev/dax device on this host: anonymous memory stands in for the CXL.mem window
    p = ::mmap(nullptr, capacity_, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (p == MAP_FAILED) return ErrorCode::OUT_OF_MEMORY;
  }
  base_ = static_cast<uint8_t*>(p);
  map_len_ = capacity_;
  if (opts_.numa_node >= 0 && opts_.numa_node < 64) {
    unsigned long mask = 1ul << opts_.numa_node;
    numa_bound_ = ::syscall(SYS_mbind, base_, map_len_, 2 /*MPOL_BIND*/, &mask, sizeof(mask) * 8, 0) == 0;
  }
  rkey_ = fnv64(pool_id_ + opts_.mount_path) & 0xFFFFFFFFull;
  init_allocator();
  initialized_ = true;
  return ErrorCode::OK;
}

void CxlMemoryBackend::shutdown() {
  if (base_) ::munmap(base_, map_len_);
  if (fd_ >= 0) ::close(fd_);
  base_ = nullptr;
  fd_ = -1;
  initialized_ = false;
}

Result<ReservationToken> CxlMemoryBackend::reserve_shard(uint64_t size, const std::string& hint) {
  // cache-line granular sizes (reference cxl_memory_backend.h:109-111)
  return StorageBackend::reserve_shard((size + kCacheLine - 1) / kCacheLine * kCacheLine, hint);
}

ErrorCode CxlMemoryBackend::write(uint64_t offset, const void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  std::memcpy(base_ + offset, data, len);
  bytes_written_ += len;
  return ErrorCode::OK;
}
ErrorCode CxlMemoryBackend::flush() {
  if (!opts_.persistent || !base_ || !dax_) return ErrorCode::OK;
  return ::msync(base_, map_len_, MS_SYNC) == 0 ? ErrorCode::OK : ErrorCode::IO_ERROR;
}
ErrorCode CxlMemoryBackend::read(uint64_t offset, void* data, uint64_t len) {
  BB_TRY(check_range(offset, len));
  std::memcpy(data, base_ + offset, len);
  bytes_read_ += len;
  return ErrorCode::OK;
}

// ================================================================ factory
static std::mutex g_pin_mu;
static HostPinHooks g_pin_hooks;
void set_host_pin_hooks(HostPinHooks h) {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  g_pin_hooks = std::move(h);
}
HostPinHooks host_pin_hooks() {
  std::lock_guard<std::mutex> lk(g_pin_mu);
  return g_pin_hooks;
}

void set_gpu_backend_factory(GpuBackendFactory f) {
  std::lock_guard<std::mutex> lk(g_factory_mu);
  g_gpu_factory = std::move(f);
}

std::unique_ptr<StorageBackend> create_storage_backend(StorageClass sc, uint64_t capacity, const BackendOptions& opts) {
  switch (sc) {
    case StorageClass::RAM_CPU:
      return std::make_unique<RamBackend>(sc, capacity, opts);
    case StorageClass::RAM_GPU: {
      std::lock_guard<std::mutex> lk(g_factory_mu);
      return g_gpu_factory ? g_gpu_factory(capacity, opts) : nullptr;
    }
    case StorageClass::NVME:
    case StorageClass::SSD:
      return std::make_unique<IoUringDiskBackend>(sc, capacity, opts);
    case StorageClass::HDD:
      return std::make_unique<MmapDiskBackend>(sc, capacity, opts);
    case StorageClass::CXL_MEMORY:
    case StorageClass::CXL_TYPE2_DEVICE:
      return std::make_unique<CxlMemoryBackend>(sc, capacity, opts);
    default:
      return nullptr;
  }
}

}  // namespace bb::worker
