#include "common/fault.h"
#include "net/tcp.h"

#include "common/audit.h"
#include "common/tenant.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sched.h>
#include <sys/epoll.h>
#include <sys/random.h>
#include <sys/eventfd.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <chrono>
#include <thread>
#include <cstring>
#include <random>

#include "common/log.h"
#include "common/sha256.h"
#include "common/trace.h"

namespace bb::net {

namespace {
std::mutex g_token_mu;
std::string g_token;
bool g_token_init = false;
}  // namespace

void set_cluster_token(const std::string& token) {
  std::lock_guard<std::mutex> lk(g_token_mu);
  g_token = token;
  g_token_init = true;
}
std::string cluster_token() {
  std::lock_guard<std::mutex> lk(g_token_mu);
  if (!g_token_init) {
    if (const char* e = std::getenv("BB_AUTH_TOKEN")) g_token = e;
    g_token_init = true;
  }
  return g_token;
}
namespace {
std::string g_token_ro;
bool g_token_ro_init = false;
}  // namespace
void set_cluster_token_ro(const std::string& token) {
  std::lock_guard<std::mutex> lk(g_token_mu);
  g_token_ro = token;
  g_token_ro_init = true;
}
std::string cluster_token_ro() {
  std::lock_guard<std::mutex> lk(g_token_mu);
  if (!g_token_ro_init) {
    if (const char* e = std::getenv("BB_AUTH_TOKEN_RO")) g_token_ro = e;
    g_token_ro_init = true;
  }
  return g_token_ro;
}
namespace {
std::atomic<int> g_encrypt{-1};  // -1: not decided yet (BB_ENCRYPT_TRANSPORT)
}
void set_transport_encryption(bool on) { g_encrypt.store(on ? 1 : 0); }
bool transport_encryption() {
  int v = g_encrypt.load();
  if (v < 0) {
    const char* e = std::getenv("BB_ENCRYPT_TRANSPORT");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
    g_encrypt.store(v);
  }
  return v == 1;
}

namespace {
constexpr size_t kNonce = 16, kMac = 32;
constexpr char kHelloMagic[] = "BBA1";
constexpr char kHelloSecure[] = "BBA2";  // as BBA1, and every frame after the handshake is sealed (tcp.h)
constexpr char kHelloRo[] = "BBR1", kHelloRoSecure[] = "BBR2";  // the same two, proving the read-only token instead
constexpr char kHelloTenant[] = "BBT1", kHelloTenantSecure[] = "BBT2";  // a tenant: the hello carries its name, the proof is keyed by its secret
void fresh_nonce(char* out) {
  size_t got = 0;
  while (got < kNonce) {
    const ssize_t r = ::getrandom(out + got, kNonce - got, 0);
    if (r > 0) got += static_cast<size_t>(r);
    else if (errno != EINTR) break;
  }
  if (got < kNonce) {  // no getrandom (ancient kernel / seccomp): fall back to the C++ entropy source
    std::random_device rd;
    for (; got < kNonce; ++got) out[got] = static_cast<char>(rd());
  }
}
// Audit trail (common/audit.h): a refused handshake or a frame that failed authentication, with whatever name was claimed.
void audit_auth_failed(const ConnPtr& c) {
  if (!audit::enabled()) return;
  const std::string who = c->hello_tenant().empty() ? std::string("unknown") : "tenant:" + c->hello_tenant();
  audit::event("auth_failed", {{"who", who}, {"peer", c->peer()}});
}
// Denials are logged, but a port scanner must not be able to fill the disk: the first few, then every 1000th.
bool log_denial() {
  static std::atomic<uint64_t> n{0};
  const uint64_t k = n.fetch_add(1, std::memory_order_relaxed);
  return k < 8 || k % 1000 == 0;
}
// HMAC(token, role || cnonce || snonce); `nonces` is cnonce followed by snonce.
std::string handshake_mac(const std::string& token, std::string_view role, std::string_view nonces) {
  std::string msg(role);
  msg.append(nonces);
  const Sha256Digest d = hmac_sha256(token, msg);
  return std::string(reinterpret_cast<const char*>(d.data()), d.size());
}
void set_nodelay(int fd) {
  int one = 1;
  ::setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
}
void set_buffers(int fd, int bytes) {
  if (bytes <= 0) return;
  ::setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &bytes, sizeof bytes);
  ::setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &bytes, sizeof bytes);
}
bool resolve(const std::string& host, uint16_t port, sockaddr_in* out) {
  std::memset(out, 0, sizeof *out);
  out->sin_family = AF_INET;
  out->sin_port = htons(port);
  std::string h = host.empty() || host == "0.0.0.0" || host == "*" ? "0.0.0.0" : host;
  if (h == "localhost") h = "127.0.0.1";
  if (::inet_pton(AF_INET, h.c_str(), &out->sin_addr) == 1) return true;
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (::getaddrinfo(h.c_str(), nullptr, &hints, &res) != 0 || !res) return false;
  out->sin_addr = reinterpret_cast<sockaddr_in*>(res->ai_addr)->sin_addr;
  ::freeaddrinfo(res);
  return true;
}
uint32_t rd32(const char* p) {
  uint32_t v;
  std::memcpy(&v, p, 4);
  return v;
}
uint64_t rd64(const char* p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;
}
}  // namespace

int tcp_listen(const std::string& host, uint16_t port, uint16_t* bound_port, std::string* err) {
  sockaddr_in addr;
  if (!resolve(host, port, &addr)) {
    if (err) *err = "cannot resolve " + host;
    return -1;
  }
  int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC | SOCK_NONBLOCK, 0);
  if (fd < 0) {
    if (err) *err = std::strerror(errno);
    return -1;
  }
  int one = 1;
  ::setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  if (::bind(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr) != 0 || ::listen(fd, 512) != 0) {
    if (err) *err = std::string("bind/listen: ") + std::strerror(errno);
    ::close(fd);
    return -1;
  }
  if (bound_port) {
    socklen_t len = sizeof addr;
    ::getsockname(fd, reinterpret_cast<sockaddr*>(&addr), &len);
    *bound_port = ntohs(addr.sin_port);
  }
  return fd;
}

int tcp_connect(const std::string& host, uint16_t port, int timeout_ms, std::string* err) {
  sockaddr_in addr;
  if (!resolve(host == "0.0.0.0" ? "127.0.0.1" : host, port, &addr)) {
    if (err) *err = "cannot resolve " + host;
    return -1;
  }
  int fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC | SOCK_NONBLOCK, 0);
  if (fd < 0) {
    if (err) *err = std::strerror(errno);
    return -1;
  }
  int rc = ::connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof addr);
  if (rc != 0 && errno == EINPROGRESS) {
    pollfd p{fd, POLLOUT, 0};
    rc = ::poll(&p, 1, timeout_ms);
    int soerr = 0;
    socklen_t len = sizeof soerr;
    if (rc == 1) ::getsockopt(fd, SOL_SOCKET, SO_ERROR, &soerr, &len);
    if (rc != 1 || soerr != 0) {
      if (err) *err = rc != 1 ? "connect timeout" : std::strerror(soerr);
      ::close(fd);
      return -1;
    }
  } else if (rc != 0) {
    if (err) *err = std::strerror(errno);
    ::close(fd);
    return -1;
  }
  int flags = ::fcntl(fd, F_GETFL, 0);
  ::fcntl(fd, F_SETFL, flags & ~O_NONBLOCK);
  set_nodelay(fd);
  return fd;
}

bool send_all(int fd, const void* data, size_t len, int timeout_ms) {
  const char* p = static_cast<const char*>(data);
  while (len) {
    ssize_t n = ::send(fd, p, len, MSG_NOSIGNAL);
    if (n > 0) {
      p += n;
      len -= static_cast<size_t>(n);
      continue;
    }
    if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
      pollfd pf{fd, POLLOUT, 0};
      if (::poll(&pf, 1, timeout_ms) <= 0) return false;
      continue;
    }
    if (n < 0 && errno == EINTR) continue;
    return false;
  }
  return true;
}

bool recv_all(int fd, void* data, size_t len, int timeout_ms) {
  char* p = static_cast<char*>(data);
  // A control-plane answer usually arrives within tens of microseconds: look for it without
  // sleeping first (a poll() sleep costs a scheduler wake-up on top of the round trip).
  static const int spin_us = [] {
    const char* e = std::getenv("BB_RPC_SPIN_US");
    return e ? std::atoi(e) : 50;
  }();
  if (spin_us > 0 && len) {
    const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us);
    do {
      ssize_t n = ::recv(fd, p, len, MSG_DONTWAIT);
      if (n > 0) {
        p += n;
        len -= static_cast<size_t>(n);
        if (!len) return true;
        continue;
      }
      if (n == 0) return false;
      if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) return false;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    } while (std::chrono::steady_clock::now() < until);
  }
  while (len) {
    pollfd pf{fd, POLLIN, 0};
    int rc = ::poll(&pf, 1, timeout_ms);
    if (rc == 0) return false;
    if (rc < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    ssize_t n = ::recv(fd, p, len, 0);
    if (n > 0) {
      p += n;
      len -= static_cast<size_t>(n);
      continue;
    }
    if (n < 0 && (errno == EAGAIN || errno == EWOULDBLOCK || errno == EINTR)) continue;
    return false;
  }
  return true;
}

// Gathered blocking send on a non-blocking or blocking socket.
static bool send_gather(int fd, const Connection::Piece* pieces, int n, int timeout_ms) {
  iovec iov[4];
  int cnt = 0;
  for (int i = 0; i < n && cnt < 4; ++i)
    if (pieces[i].len) iov[cnt++] = iovec{const_cast<void*>(pieces[i].data), pieces[i].len};
  int first = 0;
  while (first < cnt) {
    msghdr mh{};
    mh.msg_iov = iov + first;
    mh.msg_iovlen = static_cast<size_t>(cnt - first);
    ssize_t w = ::sendmsg(fd, &mh, MSG_NOSIGNAL);
    if (w < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) {
        pollfd pf{fd, POLLOUT, 0};
        if (::poll(&pf, 1, timeout_ms) <= 0) return false;
        continue;
      }
      return false;
    }
    size_t left = static_cast<size_t>(w);
    while (first < cnt && left >= iov[first].iov_len) left -= iov[first++].iov_len;
    if (first < cnt && left) {
      iov[first].iov_base = static_cast<char*>(iov[first].iov_base) + left;
      iov[first].iov_len -= left;
    }
  }
  return true;
}

// ================================================================ Connection
Connection::~Connection() {
  if (fd_ >= 0) ::close(fd_);
}

bool Connection::enable_secure(const uint8_t rx_key[kAeadKey], const uint8_t tx_key[kAeadKey]) {
  std::lock_guard<std::mutex> lk(write_mu_);
  if (!rx_.set_key(rx_key, true) || !tx_.set_key(tx_key, false)) return false;
  secure_.store(true, std::memory_order_release);
  return true;
}

// One sealed frame: [len + tag][method][id] in clear (authenticated), then ciphertext and tag.  The payload is copied
// while it is encrypted -- zero-copy replies (a pool's bytes) must not be encrypted where they live.
bool Connection::send_sealed_locked(const char* hdr, const Aead::CSpan* payload, int n, int timeout_ms) {
  size_t len = 0;
  for (int i = 0; i < n; ++i) len += payload[i].len;
  if (len > kMaxFrame) return false;
  std::string out(kFrameHeader + len + kAeadTag, '\0');
  const uint32_t wire_len = static_cast<uint32_t>(len + kAeadTag);
  std::memcpy(&out[0], &wire_len, 4);
  std::memcpy(&out[4], hdr + 4, kFrameHeader - 4);
  if (!tx_.seal(out.data(), kFrameHeader, payload, n, &out[kFrameHeader])) return false;
  return send_all(fd_, out.data(), out.size(), timeout_ms);
}

bool Connection::send(const void* data, size_t len) {
  if (closed_.load()) return false;
  std::lock_guard<std::mutex> lk(write_mu_);
  bool ok = true;
  if (secure_.load(std::memory_order_relaxed)) {  // `data` is one or more whole frames (every sender builds them with encode_frame)
    const char* p = static_cast<const char*>(data);
    while (ok && len) {
      uint32_t flen = 0;
      if (len >= kFrameHeader) std::memcpy(&flen, p, 4);
      if (len < kFrameHeader || len - kFrameHeader < flen) {
        ok = false;
        break;
      }
      const Aead::CSpan body{p + kFrameHeader, flen};
      ok = send_sealed_locked(p, &body, 1, 10000);
      p += kFrameHeader + flen;
      len -= kFrameHeader + flen;
    }
  } else {
    ok = send_all(fd_, data, len, 10000);
  }
  if (!ok) close();
  return ok;
}

bool Connection::sendv(const Piece* pieces, int n) {
  if (closed_.load()) return false;
  std::lock_guard<std::mutex> lk(write_mu_);
  bool ok;
  if (secure_.load(std::memory_order_relaxed)) {  // pieces[0] is the frame header, the rest its payload
    Aead::CSpan body[3];
    int nb = 0;
    for (int i = 1; i < n && nb < 3; ++i) body[nb++] = Aead::CSpan{pieces[i].data, pieces[i].len};
    ok = n >= 1 && n <= 4 && pieces[0].len == kFrameHeader && send_sealed_locked(static_cast<const char*>(pieces[0].data), body, nb, 30000);
  } else {
    ok = send_gather(fd_, pieces, n, 30000);
  }
  if (!ok) close();
  return ok;
}

void Connection::close() {
  bool exp = false;
  if (closed_.compare_exchange_strong(exp, true)) ::shutdown(fd_, SHUT_RDWR);
}

// ================================================================ TcpServer
TcpServer::~TcpServer() { stop(); }

ErrorCode TcpServer::start(const std::string& host, uint16_t port, int worker_threads) {
  if (running_.load()) return ErrorCode::INVALID_STATE;
  std::string err;
  listen_fd_ = tcp_listen(host, port, &port_, &err);
  if (listen_fd_ < 0) {
    BB_LOG(ERROR) << "TcpServer: cannot listen on " << host << ":" << port << ": " << err;
    return ErrorCode::NETWORK_ERROR;
  }
  epoll_fd_ = ::epoll_create1(EPOLL_CLOEXEC);
  wake_fd_ = ::eventfd(0, EFD_CLOEXEC | EFD_NONBLOCK);
  epoll_event ev{};
  ev.events = EPOLLIN | EPOLLEXCLUSIVE;
  ev.data.fd = listen_fd_;
  ::epoll_ctl(epoll_fd_, EPOLL_CTL_ADD, listen_fd_, &ev);
  ev.events = EPOLLIN;
  ev.data.fd = wake_fd_;
  ::epoll_ctl(epoll_fd_, EPOLL_CTL_ADD, wake_fd_, &ev);
  running_.store(true);
  for (int i = 0; i < std::max(1, worker_threads); ++i) workers_.emplace_back([this] { worker_loop(); });
  return ErrorCode::OK;
}

void TcpServer::stop() {
  if (!running_.exchange(false)) return;
  uint64_t one = 1;
  ssize_t ignored = ::write(wake_fd_, &one, sizeof one);
  (void)ignored;
  for (auto& w : workers_)
    if (w.joinable()) w.join();
  workers_.clear();
  std::vector<ConnPtr> all;
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& [fd, c] : conns_) all.push_back(c);
    conns_.clear();
  }
  for (auto& c : all) {
    c->close();
    on_close(c);
  }
  if (listen_fd_ >= 0) ::close(listen_fd_);
  if (epoll_fd_ >= 0) ::close(epoll_fd_);
  if (wake_fd_ >= 0) ::close(wake_fd_);
  listen_fd_ = epoll_fd_ = wake_fd_ = -1;
}

size_t TcpServer::connection_count() const {
  std::lock_guard<std::mutex> lk(mu_);
  return conns_.size();
}

void TcpServer::drop(const ConnPtr& c) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = conns_.find(c->fd());
    if (it == conns_.end() || it->second != c) return;
    conns_.erase(it);
  }
  ::epoll_ctl(epoll_fd_, EPOLL_CTL_DEL, c->fd(), nullptr);
  c->close();
  on_close(c);
}

static int env_busy_poll_us() {
  const char* e = std::getenv("BB_RPC_BUSY_POLL_US");
  return e ? std::atoi(e) : 0;
}

void TcpServer::accept_all() {
  while (true) {
    sockaddr_in peer{};
    socklen_t len = sizeof peer;
    int cfd = ::accept4(listen_fd_, reinterpret_cast<sockaddr*>(&peer), &len, SOCK_NONBLOCK | SOCK_CLOEXEC);
    if (cfd < 0) break;
    set_nodelay(cfd);
    set_buffers(cfd, sock_buf_bytes_);
    char ip[64] = {0};
    ::inet_ntop(AF_INET, &peer.sin_addr, ip, sizeof ip);
    ConnPtr c;
    {
      std::lock_guard<std::mutex> lk(mu_);
      c = std::make_shared<Connection>(cfd, next_id_++, std::string(ip) + ":" + std::to_string(ntohs(peer.sin_port)));
      conns_[cfd] = c;
    }
    on_open(c);
    epoll_event ev{};
    ev.events = EPOLLIN | EPOLLRDHUP | EPOLLONESHOT;
    ev.data.fd = cfd;
    ::epoll_ctl(epoll_fd_, EPOLL_CTL_ADD, cfd, &ev);
  }
}

// Every pool thread waits on the shared epoll set itself (EPOLLONESHOT per connection, so one
// thread owns a connection from readiness to re-arm): a request costs one thread wake-up, not a
// reactor wake-up plus a hand-off.  With busy polling the thread keeps spinning on a zero-timeout
// epoll_wait for `busy_poll_us` after its last event, which takes the scheduler out of the
// latency of back-to-back control-plane calls.
void TcpServer::worker_loop() {
  char buf[65536];
  using SteadyClock = std::chrono::steady_clock;
  const int busy_us = busy_poll_us_ >= 0 ? busy_poll_us_ : env_busy_poll_us();
  auto spin_until = SteadyClock::now();
  bool i_spin = false;
  while (running_.load()) {
    epoll_event ev{};
    // At most one thread polls at a time: several threads hammering epoll_wait(0) serialise on the
    // epoll mutex and make every request slower.
    bool spin = false;
    if (busy_us > 0) {
      if (SteadyClock::now() < spin_until) {
        int expect = 0;
        spin = i_spin || spinner_.compare_exchange_strong(expect, 1);
        i_spin = spin;
      }
      if (!spin && i_spin) {
        spinner_.store(0);
        i_spin = false;
      }
    }
    int n = ::epoll_wait(epoll_fd_, &ev, 1, spin ? 0 : 500);
    if (n < 0) {
      if (errno == EINTR) continue;
      break;
    }
    if (n == 0) {
      // yield, not pause: a runnable thread queued behind the poller on this CPU (the client, or a
      // pool thread the kernel woke for the same event) must not wait out the poller's time slice
      if (spin) ::sched_yield();
      continue;
    }
    const int fd = ev.data.fd;
    if (fd == wake_fd_) continue;  // never drained: wakes every thread; the loop condition ends them
    if (fd == listen_fd_) {
      accept_all();
      continue;
    }
    ConnPtr c;
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto it = conns_.find(fd);
      if (it == conns_.end()) continue;
      c = it->second;
    }
    c->acquire_ownership();  // pairs with release_ownership() of the thread that re-armed this connection
    bool alive = true;
    while (true) {
      ssize_t r = ::recv(c->fd(), buf, sizeof buf, 0);
      if (r > 0) {
        c->inbuf().append(buf, static_cast<size_t>(r));
        if (c->inbuf().size() > kMaxFrame + kFrameHeader + kAeadTag) {
          alive = false;
          break;
        }
        if (static_cast<size_t>(r) < sizeof buf) break;  // drained (a later arrival re-triggers after the re-arm)
        continue;
      }
      if (r == 0) {
        alive = false;
        break;
      }
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) break;
      alive = false;
      break;
    }
    // A large message is pending: grow the buffer in big steps and receive straight into it until the message is
    // complete (bulk transfers do not bounce through 64 KiB reads, buffer appends and epoll re-arms).  The buffer never
    // grows more than kBulkStep ahead of the bytes that actually arrived, whatever the frame header claims.
    if (alive && bytes_missing(c) >= sizeof buf) {
      constexpr size_t kBulkStep = 16u << 20;
      std::string& in = c->inbuf();
      size_t have = in.size();
      const auto deadline = SteadyClock::now() + std::chrono::seconds(30);
      bool stalled = false;
      while (alive && !stalled) {
        in.resize(have);
        const size_t missing = bytes_missing(c);
        if (missing == 0) break;
        in.resize(have + std::min(missing, kBulkStep));
        while (have < in.size()) {
          ssize_t r = ::recv(c->fd(), &in[have], in.size() - have, 0);
          if (r > 0) {
            have += static_cast<size_t>(r);
            continue;
          }
          if (r == 0) {
            alive = false;
            break;
          }
          if (errno == EINTR) continue;
          if (errno != EAGAIN && errno != EWOULDBLOCK) {
            alive = false;
            break;
          }
          pollfd pf{c->fd(), POLLIN, 0};
          const int rc = ::poll(&pf, 1, 200);
          if (rc < 0 && errno != EINTR) {
            alive = false;
            break;
          }
          // a sender that stalls gives the pool thread back: the connection returns to the epoll set and the rest of
          // the message is picked up when it arrives
          if (rc == 0 || !running_.load() || SteadyClock::now() > deadline) {
            stalled = true;
            break;
          }
        }
      }
      in.resize(have);
    }
    // deliver whatever arrived, even if the peer then closed
    if (!c->inbuf().empty() && !c->closed()) {
      bool ok = false;
      try {
        ok = on_data(c);
      } catch (const std::exception& e) {
        BB_LOG(ERROR) << "connection handler threw: " << e.what();
      }
      if (!ok) alive = false;
    }
    if (!alive || c->closed()) {
      drop(c);
      continue;
    }
    c->release_ownership();
    epoll_event rearm{};
    rearm.events = EPOLLIN | EPOLLRDHUP | EPOLLONESHOT;
    rearm.data.fd = c->fd();
    if (::epoll_ctl(epoll_fd_, EPOLL_CTL_MOD, c->fd(), &rearm) != 0) drop(c);
    if (busy_us > 0) spin_until = SteadyClock::now() + std::chrono::microseconds(busy_us);
  }
  if (i_spin) spinner_.store(0);
}

// ================================================================ framed RPC
std::string encode_frame(uint32_t method, uint64_t id, const std::string& payload) {
  std::string f;
  f.resize(kFrameHeader + payload.size());
  const uint32_t len = static_cast<uint32_t>(payload.size());
  std::memcpy(&f[0], &len, 4);
  std::memcpy(&f[4], &method, 4);
  std::memcpy(&f[8], &id, 8);
  if (!payload.empty()) std::memcpy(&f[kFrameHeader], payload.data(), payload.size());
  return f;
}

size_t RpcServer::bytes_missing(const ConnPtr& c) {
  const std::string& in = c->inbuf();
  if (in.size() < kFrameHeader) return 0;
  const uint32_t len = rd32(in.data());
  if (len > kMaxFrame + kAeadTag) return 0;  // on_data closes the connection
  const size_t total = static_cast<size_t>(kFrameHeader) + len;
  return total > in.size() ? total - in.size() : 0;
}

// ---------------------------------------------------------------- shared-memory channels (server side)
struct RpcServer::ShmChan {
  ShmChanHeader* hdr = nullptr;
  ConnPtr conn;                 // the TCP connection that offered it (handlers see this connection)
  uint64_t served = 0;          // last request sequence number answered (poller-private)
  size_t owner = 0;             // the one poller that serves this channel, fixed at attach time: a partition by list index
                                // would shift when another channel closes and let two pollers answer the same request
  std::atomic<bool> closed{false};
  std::mutex ov_mu;
  std::string overflow;         // a response that did not fit: collected over TCP with kShmFetchMethod
  ~ShmChan() {
    if (hdr) ::munmap(hdr, kShmChanBytes);
  }
};

RpcServer::Reply RpcServer::dispatch(const ConnPtr& c, uint32_t method, std::string_view request, uint32_t* rmethod) {
  Reply reply;
  *rmethod = method;
  if (c && c->read_only() && !ro_methods_.count(method)) {  // a read-only member asking for more than it may: refused, not hung up on
    ro_denials_.fetch_add(1, std::memory_order_relaxed);
    if (audit::enabled()) audit::event("method_denied", {{"who", "ro-member"}, {"peer", c->peer()}, {"method", std::to_string(method)}});
    *rmethod = kDeniedMarker;
    return reply;
  }
  std::shared_ptr<const Tenant> tenant;
  if (c && !c->tenant().empty()) {  // a tenant: looked up per request, so a grant that was edited or revoked takes effect on open connections
    tenant = find_tenant(c->tenant());
    if (!tenant || (!tenant->admin && !tenant_methods_.count(method))) {
      tenant_denials_.fetch_add(1, std::memory_order_relaxed);
      if (audit::enabled())
        audit::event("method_denied", {{"who", c->tenant()}, {"peer", c->peer()}, {"method", std::to_string(method)}, {"why", tenant ? "not on the tenant list" : "tenant revoked"}});
      *rmethod = kDeniedMarker;
      return reply;
    }
  }
  const std::string_view who = !c ? std::string_view("local") : !c->tenant().empty() ? std::string_view(c->tenant()) : c->read_only() ? std::string_view("ro-member") : std::string_view("member");
  audit::Scope audit_scope(who, c ? std::string_view(c->peer()) : std::string_view());  // events raised below name the caller
  TenantScope on_behalf_of(std::move(tenant));  // handlers (and the Keystone under them) see current_tenant()
  BB_TRACE_SPAN("rpc.serve", method);  // server side of every RPC (TCP and shared-memory path) on the same timeline as the client's phases
  try {
    if (const int64_t d = fault::value("delay_rpc_ms", 0); d > 0) std::this_thread::sleep_for(std::chrono::milliseconds(d));
    if (auto vit = view_handlers_.find(method); vit != view_handlers_.end()) {
      reply = vit->second(c, request);  // no copy of the request payload
    } else if (auto it = handlers_.find(method); it != handlers_.end()) {
      reply.head = it->second(c, std::string(request));
    } else {
      *rmethod = 0x7FFFFFFFu;  // unknown-method marker
    }
  } catch (const std::exception& e) {
    BB_LOG(ERROR) << "rpc handler " << method << " threw: " << e.what();
    *rmethod = 0x7FFFFFFEu;  // handler-exception marker
    reply = Reply{};
  }
  return reply;
}

// One poller per channel up to 8 (an 8-GPU box: every rank's control connection gets its own), never more than a quarter of
// the cores: a poller spins while its channels are busy.
static size_t shm_poller_count() { return std::max<size_t>(1, std::min<size_t>(8, std::thread::hardware_concurrency() / 4)); }

ErrorCode RpcServer::shm_attach(const ConnPtr& c, const std::string& path) {
  // only paths of the form /proc/<pid>/fd/<n> (a memfd of a process on this host) are accepted
  if (path.compare(0, 6, "/proc/") != 0 || path.find("/fd/") == std::string::npos || path.find("..") != std::string::npos) return ErrorCode::INVALID_PARAMETERS;
  const int fd = ::open(path.c_str(), O_RDWR | O_CLOEXEC);
  if (fd < 0) return ErrorCode::NOT_FOUND;  // another host / namespace / user: the client stays on TCP
  struct stat st{};
  if (::fstat(fd, &st) != 0 || static_cast<size_t>(st.st_size) != kShmChanBytes) {
    ::close(fd);
    return ErrorCode::INVALID_PARAMETERS;
  }
  void* m = ::mmap(nullptr, kShmChanBytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  ::close(fd);
  if (m == MAP_FAILED) return ErrorCode::OUT_OF_MEMORY;
  auto ch = std::make_shared<ShmChan>();
  ch->hdr = static_cast<ShmChanHeader*>(m);
  if (ch->hdr->magic != kShmMagic || ch->hdr->version != 1) return ErrorCode::INVALID_PARAMETERS;
  ch->conn = c;
  ch->served = ch->hdr->req_seq.load(std::memory_order_acquire);
  std::lock_guard<std::mutex> lk(shm_mu_);
  if (shm_stopped_) return ErrorCode::WORKER_NOT_READY;  // stopping: the client stays on TCP for its last requests
  for (auto& old : shm_chans_)
    if (old->conn == c) old->closed.store(true);  // a connection has at most one channel
  shm_chans_.erase(std::remove_if(shm_chans_.begin(), shm_chans_.end(), [](const auto& x) { return x->closed.load(); }), shm_chans_.end());
  ch->owner = shm_next_owner_++ % shm_poller_count();
  shm_chans_.push_back(ch);
  shm_gen_.fetch_add(1, std::memory_order_release);
  if (!shm_run_.exchange(true)) {
    const size_t n = shm_poller_count();
    for (size_t i = 0; i < n; ++i) shm_pollers_.emplace_back([this, i] { shm_poll_loop(i); });
  }
  return ErrorCode::OK;
}

std::string RpcServer::shm_take_overflow(const ConnPtr& c) {
  std::shared_ptr<ShmChan> ch;
  {
    std::lock_guard<std::mutex> lk(shm_mu_);
    for (auto& x : shm_chans_)
      if (x->conn == c) ch = x;
  }
  if (!ch) return {};
  std::lock_guard<std::mutex> lk(ch->ov_mu);
  return std::move(ch->overflow);
}

size_t RpcServer::shm_channels() const {
  std::lock_guard<std::mutex> lk(shm_mu_);
  return shm_chans_.size();
}

// Poller i serves the channels it owns.  It spins while requests keep coming and backs off to short sleeps after `busy`
// of silence (like the epoll threads' busy-poll window).
void RpcServer::shm_poll_loop(size_t idx) {
  std::vector<std::shared_ptr<ShmChan>> mine;
  uint64_t gen = ~0ull;
  auto last_active = std::chrono::steady_clock::now();
  const auto busy = std::chrono::microseconds(2000);
  while (shm_run_.load(std::memory_order_acquire)) {
    if (shm_gen_.load(std::memory_order_acquire) != gen) {
      std::lock_guard<std::mutex> lk(shm_mu_);
      gen = shm_gen_.load();
      mine.clear();
      for (const auto& ch : shm_chans_)
        if (ch->owner == idx) mine.push_back(ch);
    }
    bool any = false;
    for (auto& ch : mine) {
      if (ch->closed.load(std::memory_order_relaxed)) continue;
      ShmChanHeader* h = ch->hdr;
      const uint64_t seq = h->req_seq.load(std::memory_order_acquire);
      if (seq == ch->served) continue;
      any = true;
      const uint32_t method = h->req_method;
      const uint32_t len = std::min<uint32_t>(h->req_len, kShmReqBytes);
      const char* req = reinterpret_cast<const char*>(h) + 4096;
      uint32_t rmethod = method;
      Reply r = dispatch(ch->conn, method, std::string_view(req, len), &rmethod);
      char* resp = reinterpret_cast<char*>(h) + 4096 + kShmReqBytes;
      const size_t rlen = r.head.size() + r.ext_len;
      if (rlen > kShmRespBytes) {  // too big for the channel: park it, the client collects it over TCP
        std::lock_guard<std::mutex> lk(ch->ov_mu);
        ch->overflow = std::move(r.head);
        if (r.ext_len) ch->overflow.append(static_cast<const char*>(r.ext), r.ext_len);
        h->resp_method = kShmOverflowMarker;
        h->resp_len = 0;
      } else {
        std::memcpy(resp, r.head.data(), r.head.size());
        if (r.ext_len) std::memcpy(resp + r.head.size(), r.ext, r.ext_len);
        h->resp_method = rmethod;
        h->resp_len = static_cast<uint32_t>(rlen);
      }
      ch->served = seq;
      h->resp_seq.store(seq, std::memory_order_release);
      served_.fetch_add(1, std::memory_order_relaxed);
      shm_served_.fetch_add(1, std::memory_order_relaxed);
    }
    const auto now = std::chrono::steady_clock::now();
    if (any) last_active = now;
    else if (now - last_active > busy)  // idle: short naps at first (a request after a pause waits <= 30 us), longer ones later
      std::this_thread::sleep_for(std::chrono::microseconds(mine.empty() ? 2000 : now - last_active > std::chrono::milliseconds(200) ? 250 : 30));
    else {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
}

void RpcServer::on_close(const ConnPtr& c) {
  // A handshake that was started and then abandoned: the peer could not verify OUR proof (it holds a different secret, or
  // claimed a tenant we do not know) and hung up before sending its own.  Nothing got in, but somebody tried.
  if (!c->authed() && !c->auth_nonces().empty()) auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
  {
    std::lock_guard<std::mutex> lk(shm_mu_);
    bool changed = false;
    for (auto& ch : shm_chans_)
      if (ch->conn == c) {
        ch->closed.store(true);
        changed = true;
      }
    if (changed) {
      shm_chans_.erase(std::remove_if(shm_chans_.begin(), shm_chans_.end(), [](const auto& x) { return x->closed.load(); }), shm_chans_.end());
      shm_gen_.fetch_add(1, std::memory_order_release);
    }
  }
  if (close_hook_) close_hook_(c);
}

void RpcServer::stop_shm() {
  std::vector<std::thread> pollers;
  {
    std::lock_guard<std::mutex> lk(shm_mu_);  // an epoll thread may be inside shm_attach, growing the poller list
    shm_stopped_ = true;                      // ... and no channel is accepted from here on
    shm_run_.store(false);
    pollers.swap(shm_pollers_);
  }
  for (auto& t : pollers)
    if (t.joinable()) t.join();
  std::lock_guard<std::mutex> lk(shm_mu_);
  shm_chans_.clear();
}

void RpcServer::stop() {
  stop_shm();
  TcpServer::stop();
  std::lock_guard<std::mutex> lk(shm_mu_);
  shm_stopped_ = false;  // no epoll thread is left that could attach; a restarted server accepts channels again
}

RpcServer::~RpcServer() { stop(); }

bool RpcServer::on_data(const ConnPtr& c) {
  std::string& in = c->inbuf();
  size_t pos = 0;
  while (in.size() - pos >= kFrameHeader) {
    const uint32_t len = rd32(&in[pos]);
    if (len > kMaxFrame + (c->secure() ? kAeadTag : 0)) return false;
    if (in.size() - pos < kFrameHeader + len) break;
    const uint32_t method = rd32(&in[pos + 4]);
    const uint64_t id = rd64(&in[pos + 8]);
    const size_t body = pos + kFrameHeader;
    const size_t frame_at = pos;
    pos += kFrameHeader + len;
    uint32_t plen = len;  // payload bytes handed to the handler
    if (c->secure()) {    // sealed frame: open it in place, the clear header is the authenticated data
      const Aead::Span ct{&in[body], len >= kAeadTag ? len - kAeadTag : 0};
      if (len < kAeadTag || !c->rx().open(&in[frame_at], kFrameHeader, &ct, 1, &in[body + len - kAeadTag])) {
        if (log_denial()) BB_LOG(WARNING) << "rpc: frame from " << c->peer() << " failed authentication (altered, replayed or out of order): closing";
        auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
        return false;
      }
      plen = len - static_cast<uint32_t>(kAeadTag);
    }
    if (!c->authed()) {
      const std::string token = cluster_token();
      if (method == kAuthMethod) {
        const std::string_view msg(in.data() + body, len);
        if (token.empty() && transport_encryption()) {
          if (log_denial()) BB_LOG(ERROR) << "rpc: encrypt_transport is set but there is no cluster token to derive keys from: refusing " << c->peer();
          auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
          c->send(encode_frame(kDeniedMarker, id, std::string()));
          return false;
        }
        std::string& nonces = c->auth_nonces();
        const std::string_view magic = msg.substr(0, std::min<size_t>(4, msg.size()));
        if (nonces.empty() && msg.size() > 4 + kNonce && msg.size() <= 4 + kNonce + kMaxTenantName &&
            (magic == kHelloTenant || magic == kHelloTenantSecure)) {
          // A tenant names itself; its secret comes from this process's tenant table.  A name we do not know (or a server
          // that serves no tenants) is answered like a wrong secret -- a reply under a throw-away key, denial at the proof.
          const std::string name(msg.substr(4 + kNonce));
          std::shared_ptr<const Tenant> t = tenant_methods_.empty() ? nullptr : find_tenant(name);
          if (magic == kHelloTenant && transport_encryption()) {
            if (log_denial()) BB_LOG(WARNING) << "rpc: tenant connection from " << c->peer() << " does not encrypt but this server requires it (encrypt_transport)";
            auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
            c->send(encode_frame(kDeniedMarker, id, std::string()));
            return false;
          }
          if (magic == kHelloTenantSecure && !Aead::available()) {
            if (log_denial()) BB_LOG(WARNING) << "rpc: " << c->peer() << " asks for an encrypted connection but libcrypto is not available here";
            auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
            c->send(encode_frame(kDeniedMarker, id, std::string()));
            return false;
          }
          c->wants_secure() = magic == kHelloTenantSecure;
          c->hello_read_only() = false;
          c->hello_tenant() = name;
          nonces.assign(msg.substr(4, kNonce));
          char sn[kNonce];
          fresh_nonce(sn);
          nonces.append(sn, kNonce);
          std::string key;
          if (t) key = t->secret;
          else {
            key.assign(kNonce, '\0');
            fresh_nonce(key.data());
          }
          std::string reply(sn, kNonce);
          reply += handshake_mac(key, "bb-srv-t:" + name, nonces);
          if (!c->send(encode_frame(kAuthMethod, id, reply))) return false;
          continue;
        }
        if (token.empty() && c->hello_tenant().empty()) {  // open cluster: nothing to prove, tell the client so
          c->set_authed();
          if (!c->send(encode_frame(kAuthMethod, id, std::string()))) return false;
          continue;
        }
        const bool hello = nonces.empty() && msg.size() == 4 + kNonce;
        const bool ro_hello = hello && (magic == kHelloRo || magic == kHelloRoSecure);
        if (ro_hello && cluster_token_ro().empty()) {
          if (log_denial()) BB_LOG(WARNING) << "rpc: " << c->peer() << " presents a read-only token but this server has none (auth_token_ro)";
          auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
          c->send(encode_frame(kDeniedMarker, id, std::string()));
          return false;
        }
        if (hello && (magic == kHelloMagic || magic == kHelloRo) && transport_encryption()) {
          if (log_denial()) BB_LOG(WARNING) << "rpc: " << c->peer() << " does not encrypt but this server requires it (encrypt_transport)";
          auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
          c->send(encode_frame(kDeniedMarker, id, std::string()));
          return false;
        }
        if (hello && (magic == kHelloSecure || magic == kHelloRoSecure) && !Aead::available()) {
          if (log_denial()) BB_LOG(WARNING) << "rpc: " << c->peer() << " asks for an encrypted connection but libcrypto is not available here";
          auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
          c->send(encode_frame(kDeniedMarker, id, std::string()));
          return false;
        }
        if (hello && (magic == kHelloMagic || magic == kHelloSecure || ro_hello)) {
          c->wants_secure() = magic == kHelloSecure || magic == kHelloRoSecure;
          c->hello_read_only() = ro_hello;
          nonces.assign(msg.substr(4));
          char sn[kNonce];
          fresh_nonce(sn);
          nonces.append(sn, kNonce);
          std::string reply(sn, kNonce);
          reply += ro_hello ? handshake_mac(cluster_token_ro(), "bb-srv-ro", nonces) : handshake_mac(token, "bb-srv", nonces);
          if (!c->send(encode_frame(kAuthMethod, id, reply))) return false;
          continue;
        }
        // the proof is checked against the secret of the role the hello named, under that role's label
        const bool ro = c->hello_read_only();
        const std::string& tname = c->hello_tenant();
        std::shared_ptr<const Tenant> ten = tname.empty() || tenant_methods_.empty() ? nullptr : find_tenant(tname);
        const std::string proven = !tname.empty() ? (ten ? ten->secret : std::string()) : ro ? cluster_token_ro() : token;
        const std::string label = !tname.empty() ? "bb-cli-t:" + tname : ro ? "bb-cli-ro" : "bb-cli";
        if (nonces.size() == 2 * kNonce && msg.size() == kMac && !proven.empty() &&
            mac_equal(msg.data(), handshake_mac(proven, label, nonces).data(), kMac)) {
          if (ro) c->set_read_only();
          if (ten) {
            c->set_tenant(tname, ten->admin);
            tenant_handshakes_.fetch_add(1, std::memory_order_relaxed);
            audit::event("tenant_admitted", {{"who", tname}, {"peer", c->peer()}, {"sealed", c->wants_secure() ? "yes" : "no"}});
          }
          c->set_authed();
          if (!c->send(encode_frame(kAuthMethod, id, std::string()))) return false;  // the last clear frame
          if (c->wants_secure()) {
            uint8_t c2s[kAeadKey], s2c[kAeadKey];
            derive_key(proven, "bb-key-c2s", nonces, c2s);
            derive_key(proven, "bb-key-s2c", nonces, s2c);
            if (!c->enable_secure(c2s, s2c)) return false;
            secure_handshakes_.fetch_add(1, std::memory_order_relaxed);
          }
          nonces.clear();
          continue;
        }
        if (log_denial()) BB_LOG(WARNING) << "rpc: failed cluster-token handshake from " << c->peer();
        auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
        c->send(encode_frame(kDeniedMarker, id, std::string()));
        return false;
      }
      if (!token.empty() || transport_encryption()) {  // (encryption without a token cannot be keyed: nothing gets in)
        if (log_denial()) BB_LOG(WARNING) << "rpc: request without the cluster token from " << c->peer();
        auth_failures_.fetch_add(1, std::memory_order_relaxed), audit_auth_failed(c);
        c->send(encode_frame(kDeniedMarker, id, std::string()));
        return false;
      }
      c->set_authed();  // open cluster
    }
    Reply reply;
    uint32_t rmethod = method;
    if (method == kShmAttachMethod) {
      const ErrorCode ec = shm_attach(c, in.substr(body, plen));
      const uint32_t e = static_cast<uint32_t>(ec);
      reply.head.assign(reinterpret_cast<const char*>(&e), 4);
    } else if (method == kShmFetchMethod) {
      reply.head = shm_take_overflow(c);
    } else {
      reply = dispatch(c, method, std::string_view(in.data() + body, plen), &rmethod);
    }
    served_.fetch_add(1, std::memory_order_relaxed);
    char hdr[kFrameHeader];
    const uint32_t rlen = static_cast<uint32_t>(reply.head.size() + reply.ext_len);
    std::memcpy(hdr, &rlen, 4);
    std::memcpy(hdr + 4, &rmethod, 4);
    std::memcpy(hdr + 8, &id, 8);
    const Connection::Piece pieces[3] = {{hdr, sizeof hdr}, {reply.head.data(), reply.head.size()}, {reply.ext, reply.ext_len}};
    if (reply.head.size() + reply.ext_len > kMaxFrame || !c->sendv(pieces, 3)) return false;
  }
  if (pos) in.erase(0, pos);
  return true;
}

RpcClient::~RpcClient() { close(); }

void RpcClient::set_bulk_buffers(int bytes) {
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ >= 0) set_buffers(fd_, bytes);
}

ErrorCode RpcClient::connect(const std::string& host, uint16_t port, int timeout_ms) {
  close();
  std::string err;
  int fd = tcp_connect(host, port, timeout_ms, &err);
  if (fd < 0) {
    BB_VLOG(1) << "RpcClient: connect " << host << ":" << port << " failed: " << err;
    return ErrorCode::CONNECTION_FAILED;
  }
  std::string token = cluster_token();
  // a process that holds only the read-only token joins as a read-only member
  const bool ro = token.empty() && !cluster_token_ro().empty();
  if (ro) token = cluster_token_ro();
  // ... and one that holds no member token at all but a tenant identity presents that (common/tenant.h)
  std::string tname;
  if (token.empty()) {
    auto [n, sec] = client_tenant();
    if (!n.empty() && !sec.empty() && n.size() <= kMaxTenantName) {
      tname = n;
      token = sec;
    }
  }
  const bool as_tenant = !tname.empty();
  const std::string srv_label = as_tenant ? "bb-srv-t:" + tname : ro ? "bb-srv-ro" : "bb-srv";
  const std::string cli_label = as_tenant ? "bb-cli-t:" + tname : ro ? "bb-cli-ro" : "bb-cli";
  const bool want_secure = transport_encryption();
  secure_ = false;
  if (want_secure) {
    std::string why;
    if (token.empty()) why = "no cluster token to derive keys from";
    else Aead::available(&why);
    if (!why.empty()) {
      BB_LOG(ERROR) << "RpcClient: encrypt_transport is set but cannot be honoured: " << why;
      ::close(fd);
      return ErrorCode::ACCESS_DENIED;
    }
  }
  if (!token.empty()) {  // mutual challenge-response on the cluster token before anything else (tcp.h)
    auto exchange = [&](uint64_t id, const std::string& body, std::string* reply) -> ErrorCode {
      const std::string f = encode_frame(kAuthMethod, id, body);
      char rh[kFrameHeader];
      if (!send_all(fd, f.data(), f.size(), timeout_ms) || !recv_all(fd, rh, sizeof rh, timeout_ms)) return ErrorCode::CONNECTION_FAILED;
      const uint32_t rlen = rd32(rh), rmethod = rd32(rh + 4);
      reply->assign(rlen <= 4096 ? rlen : 0, '\0');
      if (rlen > 4096 || (rlen && !recv_all(fd, reply->data(), rlen, timeout_ms))) return ErrorCode::CONNECTION_FAILED;
      return rmethod == kAuthMethod ? ErrorCode::OK : ErrorCode::ACCESS_DENIED;
    };
    std::string nonces(kNonce, '\0'), reply;
    fresh_nonce(nonces.data());
    const char* hello = as_tenant ? (want_secure ? kHelloTenantSecure : kHelloTenant)
                        : ro      ? (want_secure ? kHelloRoSecure : kHelloRo)
                                  : (want_secure ? kHelloSecure : kHelloMagic);
    ErrorCode ec = exchange(0, std::string(hello, 4) + nonces + tname, &reply);
    if (ec == ErrorCode::OK) {
      if (reply.size() != kNonce + kMac) {
        BB_LOG(WARNING) << "RpcClient: " << host << ":" << port << " has no cluster token but this client does";
        ec = ErrorCode::ACCESS_DENIED;
      } else {
        nonces.append(reply, 0, kNonce);
        if (!mac_equal(reply.data() + kNonce, handshake_mac(token, srv_label, nonces).data(), kMac)) {
          BB_LOG(WARNING) << "RpcClient: " << host << ":" << port
                          << (as_tenant ? " does not know tenant " + tname + " (or its secret differs)" : std::string(" does not hold this cluster's token"));
          ec = ErrorCode::ACCESS_DENIED;
        } else {
          ec = exchange(1, handshake_mac(token, cli_label, nonces), &reply);
        }
      }
    }
    if (ec == ErrorCode::OK && want_secure) {  // everything after the server's acknowledgement is sealed, both ways
      uint8_t c2s[kAeadKey], s2c[kAeadKey];
      derive_key(token, "bb-key-c2s", nonces, c2s);
      derive_key(token, "bb-key-s2c", nonces, s2c);
      if (!tx_.set_key(c2s, false) || !rx_.set_key(s2c, true)) ec = ErrorCode::INTERNAL_ERROR;
      else secure_ = true;
    }
    if (ec != ErrorCode::OK) {
      ::close(fd);
      return ec;
    }
  }
  {
    std::lock_guard<std::mutex> lk(mu_);
    fd_ = fd;
    broken_ = false;
  }
  if (host == "127.0.0.1" || host == "localhost" || host == "0.0.0.0" || host == "::1") offer_shm(std::min(timeout_ms, 1000));
  return ErrorCode::OK;
}

void RpcClient::close() {
  reader_run_.store(false);
  int fd;
  {
    std::lock_guard<std::mutex> lk(mu_);
    fd = fd_;
    if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
  }
  if (reader_.joinable()) reader_.join();
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ >= 0) ::close(fd_);
  fd_ = -1;
  drop_shm();
}

void RpcClient::enable_push(std::function<void(uint32_t, const std::string&)> cb) {
  push_cb_ = std::move(cb);
  reader_run_.store(true);
  reader_ = std::thread([this] { reader_loop(); });
}

bool RpcClient::send_request(uint32_t method, uint64_t id, const void* a, size_t a_len, const void* b, size_t b_len, int timeout_ms) {
  if (a_len + b_len > kMaxFrame) return false;
  char hdr[kFrameHeader];
  const uint32_t len = static_cast<uint32_t>(a_len + b_len + (secure_ ? kAeadTag : 0));
  std::memcpy(hdr, &len, 4);
  std::memcpy(hdr + 4, &method, 4);
  std::memcpy(hdr + 8, &id, 8);
  if (!secure_) {
    const Connection::Piece pieces[3] = {{hdr, sizeof hdr}, {a, a_len}, {b, b_len}};
    return send_gather(fd_, pieces, 3, timeout_ms);
  }
  std::string out(kFrameHeader + a_len + b_len + kAeadTag, '\0');
  std::memcpy(&out[0], hdr, kFrameHeader);
  const Aead::CSpan parts[2] = {{a, a_len}, {b, b_len}};
  if (!tx_.seal(hdr, kFrameHeader, parts, 2, &out[kFrameHeader])) return false;
  return send_all(fd_, out.data(), out.size(), timeout_ms);
}

bool RpcClient::recv_frame(uint32_t* method, uint64_t* id, std::string* head, size_t head_len, void* dst, size_t dst_cap, size_t* received,
                           int hdr_timeout_ms, int body_timeout_ms, bool* idle) {
  if (idle) *idle = false;
  if (received) *received = 0;
  char hdr[kFrameHeader];
  if (!recv_all(fd_, hdr, sizeof hdr, hdr_timeout_ms)) {
    if (idle) *idle = true;  // nothing (complete) arrived in time; the caller decides whether the socket is dead
    return false;
  }
  const uint32_t len = rd32(hdr);
  *method = rd32(hdr + 4);
  if (id) *id = rd64(hdr + 8);
  const size_t overhead = secure_ ? kAeadTag : 0;
  if (len > kMaxFrame + overhead || len < overhead) return false;  // never size a buffer from an unchecked wire length
  const size_t plen = len - overhead;
  const size_t h = dst ? std::min(head_len, plen) : plen;
  const size_t rest = plen - h;
  head->assign(h, '\0');
  if (h && !recv_all(fd_, head->data(), h, body_timeout_ms)) return false;
  if (rest > dst_cap) return false;  // cannot resynchronise the stream without draining it
  if (rest && !recv_all(fd_, dst, rest, body_timeout_ms)) return false;
  if (secure_) {
    char tag[kAeadTag];
    if (!recv_all(fd_, tag, sizeof tag, body_timeout_ms)) return false;
    const Aead::Span parts[2] = {{head->data(), h}, {dst, rest}};
    if (!rx_.open(hdr, kFrameHeader, parts, 2, tag)) {
      BB_LOG(WARNING) << "RpcClient: response failed authentication (altered, replayed or out of order): closing the connection";
      return false;
    }
  }
  if (received) *received = rest;
  return true;
}

void RpcClient::reader_loop() {
  while (reader_run_.load()) {
    uint32_t method = 0;
    uint64_t id = 0;
    std::string payload;
    bool idle = false;
    if (!recv_frame(&method, &id, &payload, 0, nullptr, 0, nullptr, 500, 30000, &idle)) {
      if (!idle) break;
      // distinguish timeout (keep going) from a dead socket
      pollfd pf{fd_, POLLIN, 0};
      int rc = ::poll(&pf, 1, 0);
      if (rc > 0 && (pf.revents & (POLLHUP | POLLERR | POLLNVAL))) break;
      char probe;
      ssize_t n = ::recv(fd_, &probe, 1, MSG_PEEK | MSG_DONTWAIT);
      if (n == 0) break;
      if (n < 0 && errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) break;
      continue;
    }
    if (method & kPushFlag) {
      if (push_cb_) push_cb_(method & ~kPushFlag, payload);
    } else {
      std::lock_guard<std::mutex> lk(resp_mu_);
      if (abandoned_.erase(id) == 0) responses_[id] = std::move(payload);
      resp_cv_.notify_all();
    }
  }
  std::lock_guard<std::mutex> lk(resp_mu_);
  broken_ = true;
  resp_cv_.notify_all();
}

// ---------------------------------------------------------------- shared-memory channel (client side)
bool RpcClient::offer_shm(int timeout_ms) {
  if (const char* e = std::getenv("BB_RPC_SHM"); e && e[0] == '0') return false;
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0 || shm_ || reader_run_.load()) return false;
  const int mfd = static_cast<int>(::syscall(SYS_memfd_create, "bb-rpc-chan", 1u /* MFD_CLOEXEC */));
  if (mfd < 0) return false;
  if (::ftruncate(mfd, static_cast<off_t>(kShmChanBytes)) != 0) {
    ::close(mfd);
    return false;
  }
  void* m = ::mmap(nullptr, kShmChanBytes, PROT_READ | PROT_WRITE, MAP_SHARED, mfd, 0);
  if (m == MAP_FAILED) {
    ::close(mfd);
    return false;
  }
  auto* h = new (m) ShmChanHeader();
  h->version = 1;
  h->req_seq.store(0);
  h->resp_seq.store(0);
  h->magic = kShmMagic;
  const std::string path = "/proc/" + std::to_string(::getpid()) + "/fd/" + std::to_string(mfd);
  auto r = call_tcp_locked(kShmAttachMethod, path, timeout_ms);
  uint32_t ec = 1;
  if (r.ok() && r.value().size() == 4) std::memcpy(&ec, r.value().data(), 4);
  if (ec != 0) {  // server elsewhere, too old (unknown method), or unwilling: TCP it is
    ::munmap(m, kShmChanBytes);
    ::close(mfd);
    return false;
  }
  shm_ = h;
  shm_fd_ = mfd;
  shm_seq_ = 0;
  return true;
}

void RpcClient::drop_shm() {
  if (shm_) ::munmap(shm_, kShmChanBytes);
  if (shm_fd_ >= 0) ::close(shm_fd_);
  shm_ = nullptr;
  shm_fd_ = -1;
}

Result<std::string> RpcClient::call(uint32_t method, const std::string& request, int timeout_ms) {
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0) return ErrorCode::CLIENT_DISCONNECTED;
  if (shm_ && request.size() <= kShmReqBytes) {
    ShmChanHeader* h = shm_;
    std::memcpy(reinterpret_cast<char*>(h) + 4096, request.data(), request.size());
    h->req_method = method;
    h->req_len = static_cast<uint32_t>(request.size());
    const uint64_t seq = ++shm_seq_;
    h->req_seq.store(seq, std::memory_order_release);
    ++shm_calls_;
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    while (h->resp_seq.load(std::memory_order_acquire) != seq) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      if ((++spins & 0xFFFu) == 0) {
        const auto waited = std::chrono::steady_clock::now() - t0;
        if (waited > std::chrono::milliseconds(timeout_ms)) {
          drop_shm();  // the server is not polling this channel (any more): do not reuse it
          return ErrorCode::OPERATION_TIMEOUT;
        }
        if (waited > std::chrono::microseconds(200)) {
          // a long handler, or the server died: the TCP connection tells which
          pollfd pf{fd_, POLLIN, 0};
          if (::poll(&pf, 1, 0) > 0 && (pf.revents & (POLLHUP | POLLERR | POLLNVAL))) {
            drop_shm();
            ::close(fd_);
            fd_ = -1;
            return ErrorCode::RPC_FAILED;
          }
          char probe;
          if (::recv(fd_, &probe, 1, MSG_PEEK | MSG_DONTWAIT) == 0) {
            drop_shm();
            ::close(fd_);
            fd_ = -1;
            return ErrorCode::RPC_FAILED;
          }
          std::this_thread::yield();
        }
      }
    }
    const uint32_t rmethod = h->resp_method;
    if (rmethod == kShmOverflowMarker) return call_tcp_locked(kShmFetchMethod, std::string(), timeout_ms);
    if (rmethod == 0x7FFFFFFFu) return ErrorCode::NOT_IMPLEMENTED;
    if (rmethod == 0x7FFFFFFEu) return ErrorCode::INTERNAL_ERROR;
    if (rmethod == kDeniedMarker) return ErrorCode::ACCESS_DENIED;
    const uint32_t rlen = h->resp_len;
    if (rlen > kShmRespBytes) {  // the length lives in memory the peer writes: never read past the mapping on its say-so
      drop_shm();
      return ErrorCode::RPC_FAILED;
    }
    return std::string(reinterpret_cast<const char*>(h) + 4096 + kShmReqBytes, rlen);
  }
  return call_tcp_locked(method, request, timeout_ms);
}

namespace {
Result<std::string> marker_or(uint32_t rmethod, std::string&& payload) {
  if (rmethod == 0x7FFFFFFFu) return ErrorCode::NOT_IMPLEMENTED;
  if (rmethod == 0x7FFFFFFEu) return ErrorCode::INTERNAL_ERROR;
  if (rmethod == kDeniedMarker) return ErrorCode::ACCESS_DENIED;
  return std::move(payload);
}
}  // namespace

Result<std::string> RpcClient::call_tcp_locked(uint32_t method, const std::string& request, int timeout_ms) {
  if (fd_ < 0) return ErrorCode::CLIENT_DISCONNECTED;
  const uint64_t id = next_id_++;
  if (!send_request(method, id, request.data(), request.size(), nullptr, 0, timeout_ms)) return ErrorCode::RPC_FAILED;
  if (reader_run_.load()) {
    std::unique_lock<std::mutex> rl(resp_mu_);
    if (!resp_cv_.wait_for(rl, std::chrono::milliseconds(timeout_ms), [&] { return responses_.count(id) || broken_; })) {
      abandoned_.insert(id);  // a late response to this id is dropped by the reader instead of piling up
      return ErrorCode::OPERATION_TIMEOUT;
    }
    auto it = responses_.find(id);
    if (it == responses_.end()) return ErrorCode::RPC_FAILED;
    std::string r = std::move(it->second);
    responses_.erase(it);
    return r;
  }
  uint32_t rmethod = 0;
  std::string payload;
  if (!recv_frame(&rmethod, nullptr, &payload, 0, nullptr, 0, nullptr, timeout_ms, timeout_ms)) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::RPC_FAILED;
  }
  return marker_or(rmethod, std::move(payload));
}

Result<std::string> RpcClient::call_gather(uint32_t method, const std::string& head, const void* ext, size_t ext_len, int timeout_ms) {
  if (reader_run_.load()) {  // push mode routes responses through the reader thread: take the copying path
    std::string req = head;
    req.append(static_cast<const char*>(ext), ext_len);
    return call(method, req, timeout_ms);
  }
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0) return ErrorCode::CLIENT_DISCONNECTED;
  if (head.size() + ext_len > kMaxFrame) return ErrorCode::INVALID_PARAMETERS;
  const uint64_t id = next_id_++;
  if (!send_request(method, id, head.data(), head.size(), ext, ext_len, timeout_ms)) return ErrorCode::RPC_FAILED;
  uint32_t rmethod = 0;
  std::string payload;
  if (!recv_frame(&rmethod, nullptr, &payload, 0, nullptr, 0, nullptr, timeout_ms, timeout_ms)) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::RPC_FAILED;
  }
  return marker_or(rmethod, std::move(payload));
}

Result<std::string> RpcClient::call_scatter(uint32_t method, const std::string& request, size_t head_len, void* dst, size_t dst_cap,
                                            size_t* received, int timeout_ms) {
  if (received) *received = 0;
  if (reader_run_.load() || dst == nullptr) {
    auto r = call(method, request, timeout_ms);
    if (!r.ok()) return r.error();
    std::string& all = r.value();
    const size_t h = std::min(head_len, all.size());
    const size_t rest = all.size() - h;
    if (rest > dst_cap) return ErrorCode::BUFFER_OVERFLOW;
    if (rest) std::memcpy(dst, all.data() + h, rest);
    if (received) *received = rest;
    all.resize(h);
    return std::move(all);
  }
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ < 0) return ErrorCode::CLIENT_DISCONNECTED;
  const uint64_t id = next_id_++;
  if (!send_request(method, id, request.data(), request.size(), nullptr, 0, timeout_ms)) return ErrorCode::RPC_FAILED;
  uint32_t rmethod = 0;
  std::string head;
  if (!recv_frame(&rmethod, nullptr, &head, head_len, dst, dst_cap, received, timeout_ms, timeout_ms)) {
    ::close(fd_);
    fd_ = -1;
    return ErrorCode::RPC_FAILED;
  }
  return marker_or(rmethod, std::move(head));
}

// ================================================================ HTTP
namespace {
std::mutex g_http_mu;
std::string g_http_token;
bool g_http_token_init = false;
// "authorization: bearer <token>" among the header lines of `head` (names and the scheme are case-insensitive)
bool bearer_matches(const std::string& head, const std::string& token) {
  size_t pos = head.find("\r\n");
  while (pos != std::string::npos) {
    const size_t line = pos + 2;
    const size_t eol = head.find("\r\n", line);
    const std::string_view l(head.data() + line, (eol == std::string::npos ? head.size() : eol) - line);
    static constexpr std::string_view kName = "authorization:";
    if (l.size() > kName.size() && std::equal(kName.begin(), kName.end(), l.begin(), [](char a, char b) { return a == std::tolower(static_cast<unsigned char>(b)); })) {
      std::string_view v = l.substr(kName.size());
      while (!v.empty() && v.front() == ' ') v.remove_prefix(1);
      static constexpr std::string_view kScheme = "bearer ";
      if (v.size() > kScheme.size() && std::equal(kScheme.begin(), kScheme.end(), v.begin(), [](char a, char b) { return a == std::tolower(static_cast<unsigned char>(b)); })) {
        v.remove_prefix(kScheme.size());
        while (!v.empty() && (v.back() == ' ' || v.back() == '\t')) v.remove_suffix(1);
        // compare digests, not strings: constant time whatever the lengths
        const Sha256Digest a = hmac_sha256("bb-http", std::string(v)), b = hmac_sha256("bb-http", token);
        return mac_equal(reinterpret_cast<const char*>(a.data()), reinterpret_cast<const char*>(b.data()), a.size());
      }
    }
    pos = eol;
  }
  return false;
}
}  // namespace
void set_http_token(const std::string& token) {
  std::lock_guard<std::mutex> lk(g_http_mu);
  g_http_token = token;
  g_http_token_init = true;
}
std::string http_token() {
  std::lock_guard<std::mutex> lk(g_http_mu);
  if (!g_http_token_init) {
    if (const char* e = std::getenv("BB_HTTP_TOKEN")) g_http_token = e;
    g_http_token_init = true;
  }
  return g_http_token;
}

bool HttpServer::on_data(const ConnPtr& c) {
  std::string& in = c->inbuf();
  while (true) {
    const size_t end = in.find("\r\n\r\n");
    if (end == std::string::npos) return in.size() < 65536;
    const std::string head = in.substr(0, end);
    in.erase(0, end + 4);
    const size_t sp1 = head.find(' ');
    const size_t sp2 = sp1 == std::string::npos ? std::string::npos : head.find(' ', sp1 + 1);
    if (sp2 == std::string::npos) return false;
    const std::string method = head.substr(0, sp1);
    std::string target = head.substr(sp1 + 1, sp2 - sp1 - 1);
    std::string query;
    const size_t qm = target.find('?');
    if (qm != std::string::npos) {
      query = target.substr(qm + 1);
      target.resize(qm);
    }
    HttpResponse r;
    const std::string need = http_token();
    if (method != "GET" && method != "HEAD") {
      r.status = 405;
      r.body = "method not allowed\n";
    } else if (!need.empty() && target != "/healthz" && !bearer_matches(head, need)) {
      r.status = 401;
      r.body = "unauthorized\n";
    } else {
      auto it = routes_.find(target);
      if (it == routes_.end()) {
        r.status = 404;
        r.body = "not found\n";
      } else {
        try {
          r = it->second(target, query);
        } catch (const std::exception& e) {
          r.status = 500;
          r.body = std::string("error: ") + e.what() + "\n";
        }
      }
    }
    const char* reason = r.status == 200 ? "OK" : r.status == 404 ? "Not Found" : r.status == 405 ? "Method Not Allowed" : r.status == 503 ? "Service Unavailable" : r.status == 401 ? "Unauthorized" : "Error";
    std::string out = "HTTP/1.1 " + std::to_string(r.status) + " " + reason + (r.status == 401 ? "\r\nWWW-Authenticate: Bearer" : "") + "\r\nContent-Type: " + r.content_type +
                      "\r\nContent-Length: " + std::to_string(r.body.size()) + "\r\nConnection: keep-alive\r\n\r\n";
    if (method != "HEAD") out += r.body;
    if (!c->send(out)) return false;
  }
}

Result<std::string> http_get(const std::string& host, uint16_t port, const std::string& path, int* status, int timeout_ms) {
  std::string err;
  int fd = tcp_connect(host, port, timeout_ms, &err);
  if (fd < 0) return ErrorCode::CONNECTION_FAILED;
  const std::string bearer = http_token();
  const std::string req = "GET " + path + " HTTP/1.1\r\nHost: " + host + (bearer.empty() ? "" : "\r\nAuthorization: Bearer " + bearer) + "\r\nConnection: close\r\n\r\n";
  std::string resp;
  bool ok = send_all(fd, req.data(), req.size(), timeout_ms);
  size_t body_at = std::string::npos, content_len = std::string::npos;
  while (ok) {
    pollfd pf{fd, POLLIN, 0};
    if (::poll(&pf, 1, timeout_ms) <= 0) break;
    char buf[8192];
    ssize_t n = ::recv(fd, buf, sizeof buf, 0);
    if (n <= 0) break;
    resp.append(buf, static_cast<size_t>(n));
    if (body_at == std::string::npos) {
      const size_t e = resp.find("\r\n\r\n");
      if (e != std::string::npos) {
        body_at = e + 4;
        const size_t cl = resp.find("Content-Length:");
        if (cl != std::string::npos && cl < e) content_len = std::strtoul(resp.c_str() + cl + 15, nullptr, 10);
      }
    }
    if (body_at != std::string::npos && content_len != std::string::npos && resp.size() >= body_at + content_len) break;
  }
  ::close(fd);
  if (body_at == std::string::npos) return ErrorCode::NETWORK_ERROR;
  if (status) *status = std::atoi(resp.c_str() + 9);
  return resp.substr(body_at, content_len == std::string::npos ? std::string::npos : content_len);
}

}  // namespace bb::net
