// Network substrate: epoll reactor TCP server, framed RPC server/client, HTTP/1.1 mini server.
//
// Replaces yalantinglibs coro_rpc / coro_http (reference rpc_service.cpp:175-226, 360-413;
// unavailable offline).  A small thread pool shares one epoll set (EPOLLONESHOT per connection, so
// one thread owns a connection from readiness to re-arm): slow handlers do not block accept/IO,
// requests on one connection stay ordered, and a request costs a single thread wake-up.  Frames: [u32 len][u32 method][u64 id][payload].
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common/error.h"
#include "common/result.h"
#include "net/aead.h"

namespace bb::net {

int tcp_listen(const std::string& host, uint16_t port, uint16_t* bound_port, std::string* err = nullptr);
int tcp_connect(const std::string& host, uint16_t port, int timeout_ms, std::string* err = nullptr);
bool send_all(int fd, const void* data, size_t len, int timeout_ms = 30000);
bool recv_all(int fd, void* data, size_t len, int timeout_ms = 30000);

class TcpServer;

// Shared-secret gate for the framed RPC protocol (the reference lists mTLS / ACL as roadmap, README.md:146-153; its
// servers accept every connection).  When a cluster token is set -- `auth_token:` in keystone / worker YAML,
// BlackbirdClientOptions::auth_token, or BB_AUTH_TOKEN in the environment of every process -- an RpcServer started in
// this process answers nothing on a connection until it has passed the kAuthMethod handshake, which every RpcClient
// runs right after connecting.  The token itself never travels -- it keys an HMAC-SHA256 over fresh nonces, in both
// directions, so a passive listener learns nothing reusable and a client with a token refuses a server that cannot
// prove it has the same one:
//   C -> S  kAuthMethod  "BBA1" cnonce[16]
//   S -> C  kAuthMethod  snonce[16] HMAC(token, "bb-srv" cnonce snonce)[32]      (empty body: the server has no token)
//   C -> S  kAuthMethod  HMAC(token, "bb-cli" cnonce snonce)[32]
//   S -> C  kAuthMethod  (empty) = accepted   |   kDeniedMarker, then the server hangs up
// By default the stream after the handshake is neither encrypted nor MAC'd: the token is an admission gate.
//
// Secure mode (`encrypt_transport: true` in keystone / worker YAML, BlackbirdClientOptions::encrypt_transport,
// --encrypt-transport, or BB_ENCRYPT_TRANSPORT=1; needs a token): the client opens with "BBA2" instead of "BBA1", and once
// the handshake has passed every frame in both directions is protected with AES-256-GCM (net/aead.h):
//   [u32 len + 16][u32 method][u64 id][ciphertext(len)][tag(16)]        header = additional authenticated data
// Keys are per connection and per direction, HMAC(token, "bb-key-c2s" | "bb-key-s2c" || cnonce || snonce); the IV is the
// direction's frame counter, so a replayed, dropped, reordered or altered frame fails authentication and the connection
// is closed.  A server in secure mode refuses "BBA1" clients; a server that is not still accepts "BBA2" ones (rolling
// change).  This is a pre-shared-key channel (confidentiality + integrity for everyone who holds the token), not mTLS:
// there are no per-client identities.  Shared-memory channels (same host) carry plain frames; HTTP endpoints
// (/metrics, /healthz) stay open and clear.
//
// Read-only members (`auth_token_ro:` / BB_AUTH_TOKEN_RO on the servers; a client that holds ONLY that token -- `--auth-token-ro`,
// BlackbirdClientOptions::auth_token_ro, BB_AUTH_TOKEN_RO -- uses it): the same handshake with "BBR1" / "BBR2" and the
// role-bound labels "bb-srv-ro" / "bb-cli-ro", keyed by the second secret.  The connection is then marked read-only and the
// server dispatches only the methods its owner put on the read-only list (RpcServer::allow_read_only: object_exists /
// get_workers / stats / listings on the Keystone, D_READ / D_CHECKSUM on a data server, get / list / watch on the
// coordination store); everything else is answered with the denial marker (ACCESS_DENIED at the caller) and the
// connection stays open.  Holders of the read-only token cannot put, remove, migrate, register workers or touch leases.
//
// Tenants (common/tenant.h; a client that holds NO member token but a tenant name + secret): "BBT1" / "BBT2" cnonce[16]
// name[1..64], labels "bb-srv-t:<name>" / "bb-cli-t:<name>", keyed by that tenant's secret from the server's tenant table.
// The connection then carries the tenant's name: the server dispatches only the methods on its tenant list
// (RpcServer::allow_tenants; everything for `admin` tenants) and the handlers check keys against the tenant's grants.
// An unknown name gets the same answer as a wrong secret (a reply MAC'd with a throw-away key, then the denial), so names
// cannot be probed.
void set_cluster_token(const std::string& token);
std::string cluster_token();
void set_cluster_token_ro(const std::string& token);
std::string cluster_token_ro();
void set_transport_encryption(bool on);
bool transport_encryption();
constexpr uint32_t kAuthMethod = 0x7FFFFF00u;
// Same-host fast path of the framed RPC protocol: after connecting (and authenticating) over TCP a client may offer a
// shared-memory channel -- a memfd holding one request and one response area -- by sending its /proc/<pid>/fd/<n> path in
// a kShmAttachMethod frame.  A server on the same host (same user / PID namespace) maps it and serves the channel from a
// small pool of polling threads: a call is then a store into shared memory on both sides (~1-2 us round trip instead of
// ~15 us through the loopback TCP stack + epoll wake-up).  The TCP connection stays: it carries push frames, messages
// that do not fit the channel, and its closing detaches the channel.  BB_RPC_SHM=0 disables the offer.
// (UCX gives the reference the same thing under the hood: its `sysv,posix,cma` transports for intra-node peers,
// examples/benchmark_ucx_transports.cpp:146-149.)
constexpr uint32_t kShmAttachMethod = 0x7FFFFF01u;
constexpr uint32_t kShmFetchMethod = 0x7FFFFF02u;   // collect a response that did not fit the channel's response area
constexpr uint32_t kShmOverflowMarker = 0x7FFFFFFCu;
constexpr size_t kShmReqBytes = 2u << 20, kShmRespBytes = 6u << 20;
struct ShmChanHeader {  // at offset 0 of the mapping; request bytes at 4096, response bytes at 4096 + kShmReqBytes
  uint32_t magic;       // 'BBSH'
  uint32_t version;
  alignas(64) std::atomic<uint64_t> req_seq;   // client: ++ after the request is in place (release)
  uint32_t req_method;
  uint32_t req_len;
  alignas(64) std::atomic<uint64_t> resp_seq;  // server: = req_seq after the response is in place (release)
  uint32_t resp_method;                        // echo of the method, or one of the 0x7FFFFFFx markers
  uint32_t resp_len;
};
constexpr uint32_t kShmMagic = 0x48534242u;
constexpr size_t kShmChanBytes = 4096 + kShmReqBytes + kShmRespBytes;
constexpr uint32_t kDeniedMarker = 0x7FFFFFFDu;

class Connection : public std::enable_shared_from_this<Connection> {
 public:
  Connection(int fd, uint64_t id, std::string peer) : fd_(fd), id_(id), peer_(std::move(peer)) {}
  ~Connection();
  // Thread-safe; may be called from any thread (server push).
  bool send(const void* data, size_t len);
  bool send(const std::string& s) { return send(s.data(), s.size()); }
  // Gathered send of up to 4 pieces under one write lock (frame header + reply head + payload that stays where it is).
  struct Piece {
    const void* data;
    size_t len;
  };
  bool sendv(const Piece* pieces, int n);
  void close();
  bool closed() const { return closed_.load(); }
  uint64_t id() const { return id_; }
  const std::string& peer() const { return peer_; }
  std::string& inbuf() { return inbuf_; }
  bool authed() const { return authed_.load(std::memory_order_acquire); }
  void set_authed() { authed_.store(true, std::memory_order_release); }
  std::string& auth_nonces() { return auth_nonces_; }  // handshake in progress: cnonce + snonce
  bool& wants_secure() { return wants_secure_; }       // the client opened with "BBA2" / "BBR2"
  bool& hello_read_only() { return hello_ro_; }        // the client opened with "BBR1" / "BBR2" (handshake in progress)
  bool read_only() const { return read_only_.load(std::memory_order_acquire); }
  void set_read_only() { read_only_.store(true, std::memory_order_release); }
  // Tenant connections (common/tenant.h): the name the peer proved in the handshake ("" = a member or an open cluster).
  // Written by the owning thread before set_authed(), read by handlers afterwards.
  const std::string& tenant() const { return tenant_; }
  bool tenant_admin() const { return tenant_admin_; }
  void set_tenant(std::string name, bool admin) {
    tenant_ = std::move(name);
    tenant_admin_ = admin;
  }
  std::string& hello_tenant() { return hello_tenant_; }  // the name a "BBT1" / "BBT2" hello claimed (handshake in progress)
  // Secure mode: from now on send() / sendv() seal every frame and the owner opens incoming ones with rx().
  bool enable_secure(const uint8_t rx_key[kAeadKey], const uint8_t tx_key[kAeadKey]);
  bool secure() const { return secure_.load(std::memory_order_acquire); }
  Aead& rx() { return rx_; }
  // EPOLLONESHOT hands a connection from one pool thread to the next through the kernel; these make the
  // hand-off an explicit release/acquire pair on the connection's own state as well.
  void release_ownership() { handoff_.fetch_add(1, std::memory_order_release); }
  void acquire_ownership() { (void)handoff_.load(std::memory_order_acquire); }
  int fd() const { return fd_; }
  // arbitrary per-connection state for protocol layers (e.g. watch subscriptions)
  std::shared_ptr<void> user;

 private:
  friend class TcpServer;
  int fd_;
  uint64_t id_;
  std::string peer_;
  std::string inbuf_;
  std::mutex write_mu_;
  std::atomic<bool> closed_{false};
  std::atomic<bool> authed_{false};
  std::string auth_nonces_;
  bool wants_secure_ = false;
  bool hello_ro_ = false;
  std::string hello_tenant_, tenant_;
  bool tenant_admin_ = false;
  std::atomic<bool> read_only_{false};
  std::atomic<bool> secure_{false};
  Aead rx_, tx_;  // rx_: the thread that owns the connection; tx_: under write_mu_
  bool send_sealed_locked(const char* hdr, const Aead::CSpan* payload, int n, int timeout_ms);
  std::atomic<uint64_t> handoff_{0};
};
using ConnPtr = std::shared_ptr<Connection>;

class TcpServer {
 public:
  TcpServer() = default;
  virtual ~TcpServer();
  TcpServer(const TcpServer&) = delete;
  TcpServer& operator=(const TcpServer&) = delete;

  ErrorCode start(const std::string& host, uint16_t port, int worker_threads = 2);
  // Threads keep polling (no sleep) for this long after their last event; set before start().
  void set_busy_poll_us(int us) { busy_poll_us_ = us; }
  // Socket buffer size requested for accepted connections (0 = kernel default); set before start().
  void set_socket_buffers(int bytes) { sock_buf_bytes_ = bytes; }
  virtual void stop();
  bool running() const { return running_.load(); }
  uint16_t port() const { return port_; }
  size_t connection_count() const;

 protected:
  // Called on a worker thread with newly received bytes appended to c->inbuf(); implementations
  // consume complete messages from the buffer.  Return false to close the connection.
  virtual bool on_data(const ConnPtr& c) = 0;
  virtual void on_open(const ConnPtr&) {}
  virtual void on_close(const ConnPtr&) {}
  // Bytes still missing from the message at the front of c->inbuf() (0 = unknown / none).  When a large message is
  // pending the pool thread sizes the buffer once and receives straight into it until the message is complete
  // (bulk transfers do not bounce through 64 KiB reads and epoll re-arms).
  virtual size_t bytes_missing(const ConnPtr&) { return 0; }

 private:
  void accept_all();
  void worker_loop();
  void drop(const ConnPtr& c);

  int listen_fd_ = -1;
  int epoll_fd_ = -1;
  int wake_fd_ = -1;
  uint16_t port_ = 0;
  std::atomic<bool> running_{false};
  std::vector<std::thread> workers_;
  mutable std::mutex mu_;
  std::atomic<int> spinner_{0};
  int sock_buf_bytes_ = 0;
  int busy_poll_us_ = -1;  // -1: BB_RPC_BUSY_POLL_US from the environment (default 0 = sleep in epoll_wait)
  std::unordered_map<int, ConnPtr> conns_;
  uint64_t next_id_ = 1;
};

// ---------------------------------------------------------------- framed RPC
constexpr uint32_t kFrameHeader = 16;
constexpr uint32_t kMaxFrame = 256u << 20;
constexpr uint32_t kPushFlag = 0x80000000u;  // method field of unsolicited server -> client frames

struct Frame {
  uint32_t method = 0;
  uint64_t id = 0;
  std::string payload;
};
std::string encode_frame(uint32_t method, uint64_t id, const std::string& payload);

class RpcServer : public TcpServer {
 public:
  // Handler returns the response payload.  It runs on a worker thread.
  using Handler = std::function<std::string(const ConnPtr&, const std::string& request)>;
  void register_method(uint32_t method, Handler h) { handlers_[method] = std::move(h); }
  // Bulk methods: the request is a view into the connection buffer (no copy of the payload) and the reply may point
  // at memory that stays where it is (`ext`, e.g. a shard inside a pool) -- it is sent with a gathered write.
  struct Reply {
    std::string head;
    const void* ext = nullptr;
    size_t ext_len = 0;
  };
  using ViewHandler = std::function<Reply(const ConnPtr&, std::string_view request)>;
  void register_view_method(uint32_t method, ViewHandler h) { view_handlers_[method] = std::move(h); }
  // Methods a read-only member (a connection admitted with the read-only token) may call; set before start().
  void allow_read_only(std::initializer_list<uint32_t> methods) { ro_methods_.insert(methods.begin(), methods.end()); }
  uint64_t read_only_denials() const { return ro_denials_.load(); }
  // Methods a tenant connection may call (tenants marked `admin` may call everything); set before start().  A server that
  // allows none admits no tenants at all (their hello is refused).
  void allow_tenants(std::initializer_list<uint32_t> methods) { tenant_methods_.insert(methods.begin(), methods.end()); }
  uint64_t tenant_denials() const { return tenant_denials_.load(); }
  uint64_t tenant_handshakes() const { return tenant_handshakes_.load(); }
  void set_close_hook(std::function<void(const ConnPtr&)> f) { close_hook_ = std::move(f); }
  static bool push(const ConnPtr& c, uint32_t topic, const std::string& payload) {
    return c->send(encode_frame(kPushFlag | topic, 0, payload));
  }
  uint64_t requests_served() const { return served_.load(); }
  uint64_t shm_requests_served() const { return shm_served_.load(); }
  uint64_t secure_handshakes() const { return secure_handshakes_.load(); }  // connections that switched to sealed frames
  uint64_t auth_failures() const { return auth_failures_.load(); }          // denied handshakes + frames that failed authentication
  size_t shm_channels() const;
  ~RpcServer() override;
  void stop() override;  // channel pollers first (no request may reach a handler once stop() returned), then the sockets
  void stop_shm();       // joins the channel pollers

 protected:
  bool on_data(const ConnPtr& c) override;
  size_t bytes_missing(const ConnPtr& c) override;
  void on_close(const ConnPtr& c) override;

 private:
  struct ShmChan;
  Reply dispatch(const ConnPtr& c, uint32_t method, std::string_view request, uint32_t* rmethod);
  ErrorCode shm_attach(const ConnPtr& c, const std::string& path);
  std::string shm_take_overflow(const ConnPtr& c);
  void shm_poll_loop(size_t idx);
  mutable std::mutex shm_mu_;
  std::vector<std::shared_ptr<ShmChan>> shm_chans_;
  std::atomic<uint64_t> shm_gen_{0};
  size_t shm_next_owner_ = 0;  // guarded by shm_mu_: round-robin assignment of channels to pollers
  bool shm_stopped_ = false;   // guarded by shm_mu_: stop() is joining the pollers
  std::vector<std::thread> shm_pollers_;
  std::atomic<bool> shm_run_{false};
  std::atomic<uint64_t> shm_served_{0};
  std::atomic<uint64_t> secure_handshakes_{0}, auth_failures_{0}, ro_denials_{0}, tenant_denials_{0}, tenant_handshakes_{0};
  std::set<uint32_t> ro_methods_, tenant_methods_;
  std::unordered_map<uint32_t, Handler> handlers_;
  std::unordered_map<uint32_t, ViewHandler> view_handlers_;
  std::function<void(const ConnPtr&)> close_hook_;
  std::atomic<uint64_t> served_{0};
};

class RpcClient {
 public:
  RpcClient() = default;
  ~RpcClient();
  RpcClient(const RpcClient&) = delete;
  RpcClient& operator=(const RpcClient&) = delete;

  ErrorCode connect(const std::string& host, uint16_t port, int timeout_ms = 3000);
  // Asks the kernel for large socket buffers (bulk data connections); best effort.
  void set_bulk_buffers(int bytes = 4 << 20);
  void close();
  bool connected() const { return fd_ >= 0; }
  // connected and (push mode) the reader thread has not seen the peer go away
  bool healthy() {
    if (fd_ < 0) return false;
    std::lock_guard<std::mutex> lk(resp_mu_);
    return !broken_;
  }
  // Blocking call; thread-safe (calls are serialised per client).
  Result<std::string> call(uint32_t method, const std::string& request, int timeout_ms = 30000);
  // Bulk variants (not available in push mode).  call_gather: the request is `head` followed by `ext_len` bytes at
  // `ext`, sent with one gathered write (no copy of the payload).  call_scatter: the first `head_len` bytes of the
  // response are returned, the rest is received straight into `dst` (`*received` = its length, at most dst_cap).
  Result<std::string> call_gather(uint32_t method, const std::string& head, const void* ext, size_t ext_len, int timeout_ms = 30000);
  Result<std::string> call_scatter(uint32_t method, const std::string& request, size_t head_len, void* dst, size_t dst_cap, size_t* received,
                                   int timeout_ms = 30000);
  // Installs a handler for server push frames and starts a reader thread.  After this, call()
  // responses are also routed through the reader thread.
  void enable_push(std::function<void(uint32_t topic, const std::string& payload)> cb);
  // Offers the server a shared-memory channel (see kShmAttachMethod).  Called by connect() for loopback peers; returns
  // false (and stays on TCP) when the server is not on this host, predates the feature, or BB_RPC_SHM=0.
  bool offer_shm(int timeout_ms = 1000);
  bool shm_active() const { return shm_ != nullptr; }
  uint64_t shm_calls() const { return shm_calls_; }

 private:
  void reader_loop();
  Result<std::string> call_tcp_locked(uint32_t method, const std::string& request, int timeout_ms);
  void drop_shm();
  // One request frame out of up to two payload pieces (sealed in secure mode).  Caller holds mu_.
  bool send_request(uint32_t method, uint64_t id, const void* a, size_t a_len, const void* b, size_t b_len, int timeout_ms);
  // One frame in: the first `head_len` payload bytes into *head, the rest into dst (nullptr: everything is head).  In
  // secure mode the payload is opened in place before it is returned.  False = stream unusable (caller closes).
  bool recv_frame(uint32_t* method, uint64_t* id, std::string* head, size_t head_len, void* dst, size_t dst_cap, size_t* received,
                  int hdr_timeout_ms, int body_timeout_ms, bool* idle = nullptr);
  bool secure_ = false;
  Aead tx_, rx_;  // tx_: under mu_; rx_: the one thread that reads responses (the caller, or the push reader)
 public:
  bool secure() const { return secure_; }
 private:
  ShmChanHeader* shm_ = nullptr;  // mapped channel (nullptr = TCP only)
  int shm_fd_ = -1;
  uint64_t shm_seq_ = 0;
  uint64_t shm_calls_ = 0;
  int fd_ = -1;
  std::mutex mu_;
  uint64_t next_id_ = 1;
  // push mode
  std::function<void(uint32_t, const std::string&)> push_cb_;
  std::thread reader_;
  std::atomic<bool> reader_run_{false};
  std::mutex resp_mu_;
  std::condition_variable resp_cv_;
  std::map<uint64_t, std::string> responses_;
  std::set<uint64_t> abandoned_;  // ids whose caller timed out (guarded by resp_mu_)
  bool broken_ = false;
};

// ---------------------------------------------------------------- HTTP (GET only)
struct HttpResponse {
  int status = 200;
  std::string content_type = "text/plain; charset=utf-8";
  std::string body;
};

// The observability endpoints (/metrics, /stats) name workers, pools, tenants and their holdings.  With a token set --
// `http_auth_token:` in keystone / worker YAML, --http-token, BB_HTTP_TOKEN -- every route except /healthz (liveness probes
// carry no credentials) answers 401 unless the request has `Authorization: Bearer <token>` (what a Prometheus scrape
// config's `bearer_token` sends); http_get() of a process that holds the token sends it.
void set_http_token(const std::string& token);
std::string http_token();

class HttpServer : public TcpServer {
 public:
  using Route = std::function<HttpResponse(const std::string& path, const std::string& query)>;
  void route(const std::string& path, Route r) { routes_[path] = std::move(r); }

 protected:
  bool on_data(const ConnPtr& c) override;

 private:
  std::map<std::string, Route> routes_;
};

// Blocking HTTP GET helper (tests, CLI).
Result<std::string> http_get(const std::string& host, uint16_t port, const std::string& path, int* status = nullptr,
                             int timeout_ms = 3000);

}  // namespace bb::net
