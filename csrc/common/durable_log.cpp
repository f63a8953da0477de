#include "common/durable_log.h"

#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <sys/uio.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <vector>

#include "common/checksum.h"
#include "common/log.h"

namespace bb {

namespace {
constexpr uint32_t kSnapMagic = 0x50534242u;  // "BBSP"

bool write_all(int fd, const void* p, size_t n) {
  const char* c = static_cast<const char*>(p);
  while (n) {
    const ssize_t w = ::write(fd, c, n);
    if (w < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    c += w;
    n -= static_cast<size_t>(w);
  }
  return true;
}

bool read_file(const std::string& path, std::string* out) {
  const int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
  if (fd < 0) return false;
  struct stat st{};
  if (::fstat(fd, &st) != 0) {
    ::close(fd);
    return false;
  }
  out->resize(static_cast<size_t>(st.st_size));
  size_t got = 0;
  while (got < out->size()) {
    const ssize_t r = ::read(fd, out->data() + got, out->size() - got);
    if (r < 0 && errno == EINTR) continue;
    if (r <= 0) break;
    got += static_cast<size_t>(r);
  }
  ::close(fd);
  out->resize(got);
  return true;
}

bool mkdirs(const std::string& dir) {
  std::string cur;
  for (size_t i = 0; i <= dir.size(); ++i) {
    if (i == dir.size() || dir[i] == '/') {
      if (!cur.empty() && ::mkdir(cur.c_str(), 0755) != 0 && errno != EEXIST) return false;
    }
    if (i < dir.size()) cur.push_back(dir[i]);
  }
  return true;
}
}  // namespace

DurableLog::~DurableLog() { close(); }

std::string DurableLog::wal_path(uint64_t gen) const { return opts_.dir + "/" + opts_.name + ".wal." + std::to_string(gen); }
std::string DurableLog::snap_path(uint64_t gen) const { return opts_.dir + "/" + opts_.name + ".snap." + std::to_string(gen); }

ErrorCode DurableLog::open_gen(uint64_t gen) {
  const int fd = ::open(wal_path(gen).c_str(), O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0644);
  if (fd < 0) {
    BB_LOG(ERROR) << "durable log: cannot open " << wal_path(gen) << ": " << std::strerror(errno);
    return ErrorCode::IO_ERROR;
  }
  struct stat st{};
  ::fstat(fd, &st);
  fd_ = fd;
  gen_ = gen;
  gen_bytes_.store(static_cast<uint64_t>(st.st_size));
  if (opts_.fsync && dir_fd_ >= 0) ::fsync(dir_fd_);  // the new file's directory entry
  return ErrorCode::OK;
}

ErrorCode DurableLog::open(const Options& opts, std::string* snapshot, const std::function<void(std::string_view)>& replay) {
  close();
  opts_ = opts;
  if (opts_.dir.empty()) return ErrorCode::INVALID_PARAMETERS;
  if (!mkdirs(opts_.dir)) return ErrorCode::IO_ERROR;
  dir_fd_ = ::open(opts_.dir.c_str(), O_RDONLY | O_DIRECTORY | O_CLOEXEC);
  // ---- inventory
  std::vector<uint64_t> wals, snaps;
  if (DIR* d = ::opendir(opts_.dir.c_str())) {
    const std::string wp = opts_.name + ".wal.", sp = opts_.name + ".snap.";
    while (dirent* e = ::readdir(d)) {
      const std::string n = e->d_name;
      auto num = [&](const std::string& pfx, std::vector<uint64_t>& out) {
        if (n.compare(0, pfx.size(), pfx) != 0) return;
        const std::string t = n.substr(pfx.size());
        if (t.empty() || t.find_first_not_of("0123456789") != std::string::npos) return;  // skips *.tmp
        out.push_back(std::strtoull(t.c_str(), nullptr, 10));
      };
      num(wp, wals);
      num(sp, snaps);
    }
    ::closedir(d);
  }
  std::sort(wals.begin(), wals.end());
  std::sort(snaps.rbegin(), snaps.rend());
  // ---- newest valid snapshot
  uint64_t base_gen = 0;
  if (snapshot) snapshot->clear();
  for (uint64_t g : snaps) {
    std::string raw;
    if (!read_file(snap_path(g), &raw) || raw.size() < 16) continue;
    uint32_t magic, crc;
    uint64_t len;
    std::memcpy(&magic, raw.data(), 4);
    std::memcpy(&len, raw.data() + 4, 8);
    std::memcpy(&crc, raw.data() + 12, 4);
    if (magic != kSnapMagic || len != raw.size() - 16 || crc32c(raw.data() + 16, len) != crc) {
      BB_LOG(WARNING) << "durable log: snapshot " << snap_path(g) << " is damaged, trying an older one";
      continue;
    }
    if (snapshot) snapshot->assign(raw, 16, std::string::npos);
    base_gen = g;
    break;
  }
  // ---- replay every generation from the snapshot's on
  replayed_ = 0;
  uint64_t last_gen = base_gen;
  for (uint64_t g : wals) {
    if (g < base_gen) continue;
    last_gen = std::max(last_gen, g);
    std::string raw;
    if (!read_file(wal_path(g), &raw)) continue;
    size_t pos = 0;
    while (pos + 8 <= raw.size()) {
      uint32_t len, crc;
      std::memcpy(&len, raw.data() + pos, 4);
      std::memcpy(&crc, raw.data() + pos + 4, 4);
      if (len > raw.size() - pos - 8 || crc32c(raw.data() + pos + 8, len) != crc) break;  // torn tail
      if (replay) replay(std::string_view(raw.data() + pos + 8, len));
      ++replayed_;
      pos += 8 + len;
    }
    if (pos != raw.size()) {
      BB_LOG(WARNING) << "durable log: dropping " << (raw.size() - pos) << " torn bytes at the end of " << wal_path(g);
      if (::truncate(wal_path(g).c_str(), static_cast<off_t>(pos)) != 0) return ErrorCode::IO_ERROR;
    }
  }
  appended_ = 0;
  synced_.store(0);
  return open_gen(last_gen);
}

void DurableLog::close() {
  std::lock_guard<std::mutex> sl(sync_mu_);
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ >= 0) {
    if (opts_.fsync) ::fdatasync(fd_);
    ::close(fd_);
    fd_ = -1;
  }
  if (dir_fd_ >= 0) {
    ::close(dir_fd_);
    dir_fd_ = -1;
  }
}

uint64_t DurableLog::append(std::string_view record) {
  const uint32_t len = static_cast<uint32_t>(record.size());
  const uint32_t crc = crc32c(record.data(), record.size());
  char hdr[8];
  std::memcpy(hdr, &len, 4);
  std::memcpy(hdr + 4, &crc, 4);
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0) return 0;
  iovec iov[2] = {{hdr, 8}, {const_cast<char*>(record.data()), record.size()}};
  size_t want = 8 + record.size();
  // one writev per record: O_APPEND makes it land contiguously; retry only on a short write
  ssize_t w = ::writev(fd_, iov, 2);
  if (w < 0) return 0;
  if (static_cast<size_t>(w) != want) {
    std::string rest(hdr, 8);
    rest.append(record);
    if (!write_all(fd_, rest.data() + w, want - static_cast<size_t>(w))) return 0;
  }
  gen_bytes_.fetch_add(want, std::memory_order_relaxed);
  return ++appended_;
}

ErrorCode DurableLog::sync(uint64_t seq) {
  if (!opts_.fsync || seq == 0) return seq ? ErrorCode::OK : ErrorCode::IO_ERROR;
  if (synced_.load(std::memory_order_acquire) >= seq) return ErrorCode::OK;
  std::lock_guard<std::mutex> sl(sync_mu_);
  if (synced_.load(std::memory_order_acquire) >= seq) return ErrorCode::OK;  // a concurrent caller's fdatasync covered us
  int fd;
  uint64_t target;
  {
    std::lock_guard<std::mutex> lk(mu_);
    fd = fd_;
    target = appended_;
  }
  if (fd < 0) return ErrorCode::IO_ERROR;
  if (::fdatasync(fd) != 0) return ErrorCode::IO_ERROR;
  synced_.store(target, std::memory_order_release);
  return ErrorCode::OK;
}

uint64_t DurableLog::rotate() {
  std::lock_guard<std::mutex> sl(sync_mu_);
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0) return gen_;
  if (opts_.fsync) ::fdatasync(fd_);
  synced_.store(appended_, std::memory_order_release);
  ::close(fd_);
  fd_ = -1;
  open_gen(gen_ + 1);
  return gen_;
}

ErrorCode DurableLog::install_snapshot(uint64_t gen, std::string_view blob) {
  const std::string tmp = snap_path(gen) + ".tmp";
  const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0644);
  if (fd < 0) return ErrorCode::IO_ERROR;
  char hdr[16];
  const uint64_t len = blob.size();
  const uint32_t crc = crc32c(blob.data(), blob.size());
  std::memcpy(hdr, &kSnapMagic, 4);
  std::memcpy(hdr + 4, &len, 8);
  std::memcpy(hdr + 12, &crc, 4);
  bool ok = write_all(fd, hdr, 16) && write_all(fd, blob.data(), blob.size());
  if (ok && opts_.fsync) ok = ::fsync(fd) == 0;
  ::close(fd);
  if (!ok || ::rename(tmp.c_str(), snap_path(gen).c_str()) != 0) {
    ::unlink(tmp.c_str());
    return ErrorCode::IO_ERROR;
  }
  if (opts_.fsync && dir_fd_ >= 0) ::fsync(dir_fd_);
  // everything older than this snapshot is now redundant
  if (DIR* d = ::opendir(opts_.dir.c_str())) {
    const std::string wp = opts_.name + ".wal.", sp = opts_.name + ".snap.";
    std::vector<std::string> drop;
    while (dirent* e = ::readdir(d)) {
      const std::string n = e->d_name;
      for (const std::string* pfx : {&wp, &sp}) {
        if (n.compare(0, pfx->size(), *pfx) != 0) continue;
        const std::string t = n.substr(pfx->size());
        if (t.empty() || t.find_first_not_of("0123456789") != std::string::npos) continue;
        if (std::strtoull(t.c_str(), nullptr, 10) < gen) drop.push_back(opts_.dir + "/" + n);
      }
    }
    ::closedir(d);
    for (const auto& f : drop) ::unlink(f.c_str());
  }
  return ErrorCode::OK;
}

}  // namespace bb
