// DurableLog: append-only record log + snapshots with group commit (fdatasync), the persistence layer under the
// coordination store (`bb-coord --data-dir`) and the Keystone's local metadata log (`wal_path`).
//
// The reference keeps all of this state in memory (object table: include/blackbird/keystone/keystone_service.h:258-259;
// allocator ledger: include/blackbird/allocation/range_allocator.h:95-102) and relies on an external etcd for the
// rest; SURVEY §5.4 asks for "snapshot + log".
//
// Files in `dir`:   <name>.wal.<gen>   records appended while generation <gen> was current
//                   <name>.snap.<gen>  the owner's full state as of the instant generation <gen> began
// Recovery = newest valid snapshot S, then every record of wal.S, wal.S+1, ... in order.  A torn tail (crash in the
// middle of an append) is detected by length / CRC32C and dropped.
#pragma once
#include <atomic>
#include <cstdint>
#include <functional>
#include <mutex>
#include <string>
#include <string_view>

#include "common/error.h"

namespace bb {

class DurableLog {
 public:
  struct Options {
    std::string dir;
    std::string name = "log";
    bool fsync = true;                        // false: write(2) only (page cache) -- survives a process crash, not a power cut
    uint64_t snapshot_bytes = 64ull << 20;    // snapshot_due() once the current generation has grown past this
  };
  DurableLog() = default;
  ~DurableLog();
  DurableLog(const DurableLog&) = delete;
  DurableLog& operator=(const DurableLog&) = delete;

  // Opens (creating `dir` if needed) and replays: `snapshot` receives the newest valid snapshot blob ("" if none), then
  // `replay` is called for every record logged after it, in order.
  ErrorCode open(const Options& opts, std::string* snapshot, const std::function<void(std::string_view)>& replay);
  void close();
  bool is_open() const { return fd_ >= 0; }

  // Appends one record (buffered write).  Returns its sequence number (0 = failed).  Thread-safe; the caller's own
  // lock decides the order of records of concurrent mutations (append under that lock).
  uint64_t append(std::string_view record);
  // Group commit: returns once every record up to `seq` is on stable storage.  Concurrent callers share one fdatasync.
  ErrorCode sync(uint64_t seq);
  uint64_t append_sync(std::string_view record) {
    const uint64_t s = append(record);
    if (s) sync(s);
    return s;
  }

  bool snapshot_due() const { return gen_bytes_.load(std::memory_order_relaxed) > opts_.snapshot_bytes; }
  // Snapshot protocol.  The owner, holding the lock that orders its mutations against append(), calls rotate(): records
  // appended from now on belong to the returned generation.  It serialises its state as of that instant and passes it to
  // install_snapshot() (outside its lock): the blob is written, fsynced and renamed into place, then older files go.
  uint64_t rotate();
  ErrorCode install_snapshot(uint64_t gen, std::string_view blob);

  uint64_t generation() const { return gen_; }
  uint64_t records_replayed() const { return replayed_; }

 private:
  std::string wal_path(uint64_t gen) const;
  std::string snap_path(uint64_t gen) const;
  ErrorCode open_gen(uint64_t gen);

  Options opts_;
  int fd_ = -1;
  int dir_fd_ = -1;
  uint64_t gen_ = 0;
  std::mutex mu_;       // append / rotate
  std::mutex sync_mu_;  // one fdatasync at a time
  uint64_t appended_ = 0;                  // guarded by mu_
  std::atomic<uint64_t> synced_{0};
  std::atomic<uint64_t> gen_bytes_{0};
  uint64_t replayed_ = 0;
};

}  // namespace bb
