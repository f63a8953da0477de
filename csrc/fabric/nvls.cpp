#include "fabric/nvls.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstring>

#include "common/log.h"

namespace bb::gpu {

// ================================================================ driver entry points (no libcuda link dependency)
namespace {
template <typename F>
F load_driver(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return reinterpret_cast<F>(fn);
}
#define BB_DRV(name) static auto p_##name = load_driver<decltype(&name)>(#name)

struct Driver {
  decltype(&cuMemCreate) MemCreate = load_driver<decltype(&cuMemCreate)>("cuMemCreate");
  decltype(&cuMemRelease) MemRelease = load_driver<decltype(&cuMemRelease)>("cuMemRelease");
  decltype(&cuMemAddressReserve) MemAddressReserve = load_driver<decltype(&cuMemAddressReserve)>("cuMemAddressReserve");
  decltype(&cuMemAddressFree) MemAddressFree = load_driver<decltype(&cuMemAddressFree)>("cuMemAddressFree");
  decltype(&cuMemMap) MemMap = load_driver<decltype(&cuMemMap)>("cuMemMap");
  decltype(&cuMemUnmap) MemUnmap = load_driver<decltype(&cuMemUnmap)>("cuMemUnmap");
  decltype(&cuMemSetAccess) MemSetAccess = load_driver<decltype(&cuMemSetAccess)>("cuMemSetAccess");
  decltype(&cuMemExportToShareableHandle) MemExport = load_driver<decltype(&cuMemExportToShareableHandle)>("cuMemExportToShareableHandle");
  decltype(&cuMemImportFromShareableHandle) MemImport = load_driver<decltype(&cuMemImportFromShareableHandle)>("cuMemImportFromShareableHandle");
  decltype(&cuMemGetAllocationGranularity) MemGran = load_driver<decltype(&cuMemGetAllocationGranularity)>("cuMemGetAllocationGranularity");
  decltype(&cuMulticastCreate) McCreate = load_driver<decltype(&cuMulticastCreate)>("cuMulticastCreate");
  decltype(&cuMulticastAddDevice) McAddDevice = load_driver<decltype(&cuMulticastAddDevice)>("cuMulticastAddDevice");
  decltype(&cuMulticastBindMem) McBindMem = load_driver<decltype(&cuMulticastBindMem)>("cuMulticastBindMem");
  decltype(&cuMulticastGetGranularity) McGran = load_driver<decltype(&cuMulticastGetGranularity)>("cuMulticastGetGranularity");
  decltype(&cuMulticastUnbind) McUnbind = load_driver<decltype(&cuMulticastUnbind)>("cuMulticastUnbind");
  decltype(&cuDeviceGetAttribute) DevAttr = load_driver<decltype(&cuDeviceGetAttribute)>("cuDeviceGetAttribute");
  decltype(&cuDeviceGet) DevGet = load_driver<decltype(&cuDeviceGet)>("cuDeviceGet");
  bool ok() const {
    return MemCreate && MemRelease && MemAddressReserve && MemMap && MemSetAccess && MemExport && MemImport && MemGran && McCreate &&
           McAddDevice && McBindMem && McGran && DevAttr && DevGet;
  }
};
Driver& drv() {
  static Driver d;
  return d;
}

std::string sock_name(const std::string& tag, int rank) { return "bb-fab-" + tag + "-" + std::to_string(rank); }

int make_addr(const std::string& name, sockaddr_un* addr) {
  std::memset(addr, 0, sizeof *addr);
  addr->sun_family = AF_UNIX;
  const size_t n = std::min(name.size(), sizeof(addr->sun_path) - 2);
  std::memcpy(addr->sun_path + 1, name.data(), n);  // abstract namespace: leading NUL
  return static_cast<int>(offsetof(sockaddr_un, sun_path) + 1 + n);
}
}  // namespace

// ================================================================ FdChannel
FdChannel::FdChannel(std::string tag, int rank) : tag_(std::move(tag)), rank_(rank) {}

FdChannel::~FdChannel() {
  run_.store(false);
  if (listen_fd_ >= 0) ::shutdown(listen_fd_, SHUT_RDWR);
  if (thread_.joinable()) thread_.join();
  if (listen_fd_ >= 0) ::close(listen_fd_);
  for (auto& [n, fd] : fds_) ::close(fd);
}

ErrorCode FdChannel::start() {
  listen_fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (listen_fd_ < 0) return ErrorCode::NETWORK_ERROR;
  sockaddr_un addr;
  const int len = make_addr(sock_name(tag_, rank_), &addr);
  if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), static_cast<socklen_t>(len)) != 0 || ::listen(listen_fd_, 64) != 0) {
    BB_LOG(ERROR) << "FdChannel: cannot bind abstract socket " << sock_name(tag_, rank_) << ": " << std::strerror(errno);
    return ErrorCode::NETWORK_ERROR;
  }
  run_.store(true);
  thread_ = std::thread([this] { serve(); });
  return ErrorCode::OK;
}

void FdChannel::publish(const std::string& name, int fd) {
  std::lock_guard<std::mutex> lk(mu_);
  fds_[name] = fd;
}

void FdChannel::serve() {
  while (run_.load()) {
    pollfd pf{listen_fd_, POLLIN, 0};
    if (::poll(&pf, 1, 200) <= 0) continue;
    int c = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
    if (c < 0) continue;
    char buf[160];
    size_t got = 0;
    while (got < sizeof buf - 1) {
      ssize_t n = ::recv(c, buf + got, sizeof buf - 1 - got, 0);
      if (n <= 0) break;
      got += static_cast<size_t>(n);
      if (std::memchr(buf, '\n', got)) break;
    }
    buf[got] = 0;
    if (char* nl = std::strchr(buf, '\n')) *nl = 0;
    int fd = -1;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = fds_.find(buf);
      if (it != fds_.end()) fd = it->second;
    }
    char status = fd >= 0 ? 1 : 0;
    iovec iov{&status, 1};
    msghdr msg{};
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
    if (fd >= 0) {
      msg.msg_control = ctrl;
      msg.msg_controllen = sizeof ctrl;
      cmsghdr* cm = CMSG_FIRSTHDR(&msg);
      cm->cmsg_level = SOL_SOCKET;
      cm->cmsg_type = SCM_RIGHTS;
      cm->cmsg_len = CMSG_LEN(sizeof(int));
      std::memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    }
    ::sendmsg(c, &msg, MSG_NOSIGNAL);
    ::close(c);
  }
}

Result<int> FdChannel::fetch(const std::string& tag, int rank, const std::string& name, int timeout_ms) {
  const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
  while (std::chrono::steady_clock::now() < deadline) {
    int s = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (s < 0) return ErrorCode::NETWORK_ERROR;
    sockaddr_un addr;
    const int len = make_addr(sock_name(tag, rank), &addr);
    if (::connect(s, reinterpret_cast<sockaddr*>(&addr), static_cast<socklen_t>(len)) == 0) {
      const std::string req = name + "\n";
      if (::send(s, req.data(), req.size(), MSG_NOSIGNAL) == static_cast<ssize_t>(req.size())) {
        char status = 0;
        iovec iov{&status, 1};
        msghdr msg{};
        msg.msg_iov = &iov;
        msg.msg_iovlen = 1;
        alignas(cmsghdr) char ctrl[CMSG_SPACE(sizeof(int))];
        msg.msg_control = ctrl;
        msg.msg_controllen = sizeof ctrl;
        pollfd pf{s, POLLIN, 0};
        if (::poll(&pf, 1, 2000) > 0 && ::recvmsg(s, &msg, MSG_CMSG_CLOEXEC) == 1 && status == 1) {
          for (cmsghdr* cm = CMSG_FIRSTHDR(&msg); cm; cm = CMSG_NXTHDR(&msg, cm))
            if (cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS) {
              int fd;
              std::memcpy(&fd, CMSG_DATA(cm), sizeof(int));
              ::close(s);
              return fd;
            }
        }
      }
    }
    ::close(s);
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
  return ErrorCode::OPERATION_TIMEOUT;
}

// ================================================================ NvlsArena
bool NvlsArena::supported(int device) {
  if (!drv().ok()) return false;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device >= n) {
    cudaGetLastError();
    return false;
  }
  cudaSetDevice(device);
  cudaFree(nullptr);
  CUdevice dev;
  if (drv().DevGet(&dev, device) != CUDA_SUCCESS) return false;
  int mc = 0, fdok = 0;
  drv().DevAttr(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  drv().DevAttr(&fdok, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  return mc && fdok;
}

NvlsArena::NvlsArena(int device, int rank, int world, std::string tag, std::vector<std::vector<int>> groups, uint64_t arena_bytes)
    : device_(device), rank_(rank), world_(world), tag_(std::move(tag)), groups_(std::move(groups)), bytes_(arena_bytes) {
  const size_t n = groups_.size();
  phys_.assign(n, 0);
  mc_.assign(n, 0);
  mc_va_.assign(n, nullptr);
  local_va_.assign(n, nullptr);
  peer_va_.resize(n);
}

NvlsArena::~NvlsArena() {
  if (!drv().ok()) return;
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (size_t g = 0; g < groups_.size(); ++g) {
    for (auto& [r, va] : peer_va_[g])
      if (va && va != local_va_[g]) {
        drv().MemUnmap(reinterpret_cast<CUdeviceptr>(va), bytes_);
        drv().MemAddressFree(reinterpret_cast<CUdeviceptr>(va), bytes_);
      }
    if (mc_va_[g]) {
      drv().MemUnmap(reinterpret_cast<CUdeviceptr>(mc_va_[g]), bytes_);
      drv().MemAddressFree(reinterpret_cast<CUdeviceptr>(mc_va_[g]), bytes_);
    }
    if (local_va_[g]) {
      drv().MemUnmap(reinterpret_cast<CUdeviceptr>(local_va_[g]), bytes_);
      drv().MemAddressFree(reinterpret_cast<CUdeviceptr>(local_va_[g]), bytes_);
    }
    if (mc_[g]) drv().MemRelease(mc_[g]);
    if (phys_[g]) drv().MemRelease(phys_[g]);
  }
}

int NvlsArena::local_index(size_t g) const {
  const auto& m = groups_[g];
  auto it = std::find(m.begin(), m.end(), rank_);
  return it == m.end() ? -1 : static_cast<int>(it - m.begin());
}

ErrorCode NvlsArena::fail(const std::string& what, int r) {
  err_ = what + " failed with CUresult " + std::to_string(r);
  BB_LOG(ERROR) << "NvlsArena(rank " << rank_ << "): " << err_;
  return ErrorCode::FABRIC_ERROR;
}

void* NvlsArena::peer_ptr(size_t g, int rank) const {
  if (g >= peer_va_.size()) return nullptr;
  auto it = peer_va_[g].find(rank);
  return it == peer_va_[g].end() ? nullptr : it->second;
}

bool NvlsArena::parse_pool_id(const std::string& id, size_t* g, int* rank) {
  unsigned long gg = 0;
  int rr = 0;
  if (std::sscanf(id.c_str(), "mc%lu@gpu%d", &gg, &rr) != 2) return false;
  *g = gg;
  *rank = rr;
  return true;
}

namespace {
CUmemAllocationProp mem_prop(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}
}  // namespace

ErrorCode NvlsArena::phase1_create() {
  if (!drv().ok()) {
    err_ = "CUDA driver VMM / multicast entry points unavailable";
    return ErrorCode::NOT_IMPLEMENTED;
  }
  if (cudaSetDevice(device_) != cudaSuccess) return ErrorCode::FABRIC_ERROR;
  cudaFree(nullptr);
  chan_ = std::make_unique<FdChannel>(tag_, rank_);
  BB_TRY(chan_->start());
  CUmemAllocationProp prop = mem_prop(device_);
  size_t g_mem = 0, g_mc = 0;
  CUresult r = drv().MemGran(&g_mem, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS) return fail("cuMemGetAllocationGranularity", r);
  size_t max_members = 1;
  for (const auto& m : groups_) max_members = std::max(max_members, m.size());
  CUmulticastObjectProp mp{};
  mp.numDevices = static_cast<unsigned>(max_members);
  mp.size = bytes_;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  r = drv().McGran(&g_mc, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS) return fail("cuMulticastGetGranularity", r);
  gran_ = std::max<uint64_t>(g_mem, g_mc);
  bytes_ = (bytes_ + gran_ - 1) / gran_ * gran_;
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (size_t g = 0; g < groups_.size(); ++g) {
    if (local_index(g) < 0) continue;
    CUmemGenericAllocationHandle h;
    r = drv().MemCreate(&h, bytes_, &prop, 0);
    if (r != CUDA_SUCCESS) return fail("cuMemCreate", r);
    phys_[g] = h;
    int fd = -1;
    r = drv().MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) return fail("cuMemExportToShareableHandle(phys)", r);
    chan_->publish("phys-g" + std::to_string(g), fd);
    CUdeviceptr va = 0;
    r = drv().MemAddressReserve(&va, bytes_, gran_, 0, 0);
    if (r != CUDA_SUCCESS) return fail("cuMemAddressReserve", r);
    r = drv().MemMap(va, bytes_, 0, h, 0);
    if (r != CUDA_SUCCESS) return fail("cuMemMap(local)", r);
    r = drv().MemSetAccess(va, bytes_, &acc, 1);
    if (r != CUDA_SUCCESS) return fail("cuMemSetAccess(local)", r);
    local_va_[g] = reinterpret_cast<void*>(va);
    peer_va_[g][rank_] = local_va_[g];
    cudaMemset(local_va_[g], 0, bytes_);
    if (groups_[g].front() == rank_) {  // group leader creates the multicast object
      CUmulticastObjectProp gp{};
      gp.numDevices = static_cast<unsigned>(groups_[g].size());
      gp.size = bytes_;
      gp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
      CUmemGenericAllocationHandle mch;
      r = drv().McCreate(&mch, &gp);
      if (r != CUDA_SUCCESS) return fail("cuMulticastCreate", r);
      mc_[g] = mch;
      int mfd = -1;
      r = drv().MemExport(&mfd, mch, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
      if (r != CUDA_SUCCESS) return fail("cuMemExportToShareableHandle(mc)", r);
      chan_->publish("mc-g" + std::to_string(g), mfd);
    }
  }
  cudaDeviceSynchronize();
  return ErrorCode::OK;
}

ErrorCode NvlsArena::phase2_join() {
  cudaSetDevice(device_);
  CUdevice dev;
  if (drv().DevGet(&dev, device_) != CUDA_SUCCESS) return ErrorCode::FABRIC_ERROR;
  for (size_t g = 0; g < groups_.size(); ++g) {
    if (local_index(g) < 0) continue;
    if (!mc_[g]) {
      auto fd = FdChannel::fetch(tag_, groups_[g].front(), "mc-g" + std::to_string(g));
      if (!fd.ok()) {
        err_ = "timeout fetching the multicast handle of group " + std::to_string(g);
        return fd.error();
      }
      CUmemGenericAllocationHandle mch;
      CUresult r = drv().MemImport(&mch, reinterpret_cast<void*>(static_cast<uintptr_t>(fd.value())), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      ::close(fd.value());
      if (r != CUDA_SUCCESS) return fail("cuMemImportFromShareableHandle(mc)", r);
      mc_[g] = mch;
    }
    CUresult r = drv().McAddDevice(mc_[g], dev);
    if (r != CUDA_SUCCESS) return fail("cuMulticastAddDevice", r);
  }
  return ErrorCode::OK;
}

ErrorCode NvlsArena::phase3_bind() {
  cudaSetDevice(device_);
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (size_t g = 0; g < groups_.size(); ++g) {
    if (local_index(g) < 0) continue;
    CUresult r = drv().McBindMem(mc_[g], 0, phys_[g], 0, bytes_, 0);
    if (r != CUDA_SUCCESS) return fail("cuMulticastBindMem", r);
    CUdeviceptr va = 0;
    r = drv().MemAddressReserve(&va, bytes_, gran_, 0, 0);
    if (r != CUDA_SUCCESS) return fail("cuMemAddressReserve(mc)", r);
    r = drv().MemMap(va, bytes_, 0, mc_[g], 0);
    if (r != CUDA_SUCCESS) return fail("cuMemMap(mc)", r);
    r = drv().MemSetAccess(va, bytes_, &acc, 1);
    if (r != CUDA_SUCCESS) return fail("cuMemSetAccess(mc)", r);
    mc_va_[g] = reinterpret_cast<void*>(va);
  }
  return ErrorCode::OK;
}

ErrorCode NvlsArena::phase4_map_peers() {
  cudaSetDevice(device_);
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  for (size_t g = 0; g < groups_.size(); ++g) {
    for (int r : groups_[g]) {
      if (r == rank_) continue;
      auto fd = FdChannel::fetch(tag_, r, "phys-g" + std::to_string(g));
      if (!fd.ok()) {
        err_ = "timeout fetching arena of rank " + std::to_string(r);
        return fd.error();
      }
      CUmemGenericAllocationHandle h;
      CUresult cr = drv().MemImport(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd.value())), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      ::close(fd.value());
      if (cr != CUDA_SUCCESS) return fail("cuMemImportFromShareableHandle(phys)", cr);
      CUdeviceptr va = 0;
      cr = drv().MemAddressReserve(&va, bytes_, gran_, 0, 0);
      if (cr != CUDA_SUCCESS) return fail("cuMemAddressReserve(peer)", cr);
      cr = drv().MemMap(va, bytes_, 0, h, 0);
      if (cr != CUDA_SUCCESS) return fail("cuMemMap(peer)", cr);
      cr = drv().MemSetAccess(va, bytes_, &acc, 1);
      if (cr != CUDA_SUCCESS) return fail("cuMemSetAccess(peer)", cr);
      drv().MemRelease(h);  // the mapping keeps the allocation alive
      peer_va_[g][r] = reinterpret_cast<void*>(va);
    }
  }
  return ErrorCode::OK;
}

}  // namespace bb::gpu
