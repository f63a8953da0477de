// NVLS replica arenas: VMM slabs bound to NVSwitch multicast objects (hard part #2).
//
// The reference writes every replica from the client, one UCX put per copy
// (blackbird_client.cpp:254-261): N x the egress.  Here a replica set is a *multicast group*:
// every member GPU contributes one physical VMM allocation bound at offset 0 of a
// cuMulticast object, so a single `multimem.st` from the put kernel lands at the same offset on
// all members (1 x egress; the switch replicates).  The placement engine's symmetric-offset mode
// (alloc/allocator.h) hands out identical offsets on all members of one group.
//
// Processes exchange VMM / multicast handles as POSIX file descriptors over AF_UNIX sockets
// (SCM_RIGHTS); ranks synchronise the create -> add-device -> bind -> map phases with barriers
// supplied by the caller (torch.distributed in the Python layer).
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common/error.h"
#include "common/result.h"

namespace bb::gpu {

// Tiny fd server: peers connect to "\0bb-fab-<tag>-<rank>", send a name, receive the fd.
class FdChannel {
 public:
  FdChannel(std::string tag, int rank);
  ~FdChannel();
  ErrorCode start();
  void publish(const std::string& name, int fd);
  // Blocks (with retries) until `rank` has published `name`; returns a new fd owned by the caller.
  static Result<int> fetch(const std::string& tag, int rank, const std::string& name, int timeout_ms = 20000);

 private:
  void serve();
  std::string tag_;
  int rank_;
  int listen_fd_ = -1;
  std::atomic<bool> run_{false};
  std::thread thread_;
  std::mutex mu_;
  std::map<std::string, int> fds_;
};

class NvlsArena {
 public:
  // groups[g] = ranks that replicate together; every rank must pass the same `groups`.
  NvlsArena(int device, int rank, int world, std::string tag, std::vector<std::vector<int>> groups, uint64_t arena_bytes);
  ~NvlsArena();
  static bool supported(int device);

  // Collective bring-up in four phases; the caller runs a cross-rank barrier between phases.
  ErrorCode phase1_create();   // local physical memory + (leaders) multicast objects, publish fds
  ErrorCode phase2_join();     // import multicast objects, add this device
  ErrorCode phase3_bind();     // bind local memory, map the multicast window
  ErrorCode phase4_map_peers();// import + map every member's memory (unicast reads)

  size_t num_groups() const { return groups_.size(); }
  int rank() const { return rank_; }
  const std::vector<int>& members(size_t g) const { return groups_[g]; }
  bool member_of(size_t g) const { return local_index(g) >= 0; }
  uint64_t arena_bytes() const { return bytes_; }
  // Multicast window of group g (valid on members only); stores must use multimem.st.
  void* mc_ptr(size_t g) const { return mc_va_[g]; }
  // Unicast pointer to member `rank`'s arena of group g (valid everywhere after phase 4).
  void* peer_ptr(size_t g, int rank) const;
  static std::string pool_id(size_t g, int rank) { return "mc" + std::to_string(g) + "@gpu" + std::to_string(rank); }
  static std::string domain(size_t g) { return "nvls-g" + std::to_string(g); }
  // Parses a pool id produced by pool_id(); false when it is not an arena pool.
  static bool parse_pool_id(const std::string& id, size_t* g, int* rank);
  const std::string& last_error() const { return err_; }

 private:
  int local_index(size_t g) const;
  ErrorCode fail(const std::string& what, int curesult);

  int device_, rank_, world_;
  std::string tag_;
  std::vector<std::vector<int>> groups_;
  uint64_t bytes_ = 0;
  uint64_t gran_ = 0;
  std::unique_ptr<FdChannel> chan_;
  std::vector<unsigned long long> phys_;  // CUmemGenericAllocationHandle per group (0 = not a member)
  std::vector<unsigned long long> mc_;    // multicast handle per group
  std::vector<void*> mc_va_;
  std::vector<void*> local_va_;
  std::vector<std::map<int, void*>> peer_va_;  // [group][rank] -> unicast VA
  std::string err_;
};

}  // namespace bb::gpu
