#include "coord/etcd_coord.h"

#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <condition_variable>
#include <cstring>
#include <sstream>

#include "common/log.h"
#include "net/tcp.h"

namespace bb::coord {

// ================================================================ base64
std::string b64_encode(std::string_view raw) {
  static const char* A = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  std::string out;
  out.reserve((raw.size() + 2) / 3 * 4);
  size_t i = 0;
  for (; i + 2 < raw.size(); i += 3) {
    const uint32_t v = (static_cast<uint8_t>(raw[i]) << 16) | (static_cast<uint8_t>(raw[i + 1]) << 8) | static_cast<uint8_t>(raw[i + 2]);
    out += A[v >> 18], out += A[(v >> 12) & 63], out += A[(v >> 6) & 63], out += A[v & 63];
  }
  if (i + 1 == raw.size()) {
    const uint32_t v = static_cast<uint8_t>(raw[i]) << 16;
    out += A[v >> 18], out += A[(v >> 12) & 63], out += "==";
  } else if (i + 2 == raw.size()) {
    const uint32_t v = (static_cast<uint8_t>(raw[i]) << 16) | (static_cast<uint8_t>(raw[i + 1]) << 8);
    out += A[v >> 18], out += A[(v >> 12) & 63], out += A[(v >> 6) & 63], out += '=';
  }
  return out;
}

bool b64_decode(std::string_view text, std::string* out) {
  out->clear();
  uint32_t acc = 0;
  int bits = 0;
  for (char c : text) {
    int v;
    if (c >= 'A' && c <= 'Z') v = c - 'A';
    else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
    else if (c >= '0' && c <= '9') v = c - '0' + 52;
    else if (c == '+' || c == '-') v = 62;
    else if (c == '/' || c == '_') v = 63;
    else if (c == '=' || c == '\n' || c == '\r') continue;
    else return false;
    acc = (acc << 6) | static_cast<uint32_t>(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out->push_back(static_cast<char>((acc >> bits) & 0xFF));
    }
  }
  return true;
}

namespace {

std::string prefix_end(const std::string& prefix) {  // smallest key greater than every key with this prefix
  std::string e = prefix;
  while (!e.empty()) {
    if (static_cast<uint8_t>(e.back()) != 0xFF) {
      e.back() = static_cast<char>(static_cast<uint8_t>(e.back()) + 1);
      return e;
    }
    e.pop_back();
  }
  return std::string(1, '\0');  // "\0" = to the end of the key space
}

int64_t num(const Json& j) {  // etcd renders 64-bit integers as strings
  if (j.is_string()) return std::strtoll(j.as_string().c_str(), nullptr, 10);
  return j.as_int(0);
}

KeyValue kv_from(const Json& j) {
  KeyValue kv;
  b64_decode(j.at("key").as_string(), &kv.key);
  b64_decode(j.at("value").as_string(), &kv.value);
  kv.create_revision = num(j.at("create_revision"));
  kv.mod_revision = num(j.at("mod_revision"));
  kv.lease = num(j.at("lease"));
  return kv;
}

Json put_op(const std::string& key, const std::string& value, LeaseId lease) {
  Json p = Json::object();
  p["key"] = b64_encode(key);
  p["value"] = b64_encode(value);
  if (lease) p["lease"] = std::to_string(lease);
  return p;
}

// Reads one HTTP response head from `buf` (+ fd); returns false on a dead connection.  *chunked / *content_length
// describe the body; the head is consumed from `buf`.
bool read_head(int fd, std::string& buf, int* status, bool* chunked, int64_t* content_length, int timeout_ms) {
  size_t end;
  while ((end = buf.find("\r\n\r\n")) == std::string::npos) {
    pollfd pf{fd, POLLIN, 0};
    if (::poll(&pf, 1, timeout_ms) <= 0) return false;
    char tmp[8192];
    const ssize_t n = ::recv(fd, tmp, sizeof tmp, 0);
    if (n <= 0) return false;
    buf.append(tmp, static_cast<size_t>(n));
    if (buf.size() > (1u << 20)) return false;
  }
  std::string head = buf.substr(0, end);
  buf.erase(0, end + 4);
  *status = head.size() > 12 ? std::atoi(head.c_str() + 9) : 0;
  for (char& c : head) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
  *chunked = head.find("transfer-encoding: chunked") != std::string::npos;
  *content_length = -1;
  const size_t cl = head.find("content-length:");
  if (cl != std::string::npos) *content_length = std::strtoll(head.c_str() + cl + 15, nullptr, 10);
  return true;
}

bool fill(int fd, std::string& buf, size_t want, int timeout_ms) {
  while (buf.size() < want) {
    pollfd pf{fd, POLLIN, 0};
    if (::poll(&pf, 1, timeout_ms) <= 0) return false;
    char tmp[16384];
    const ssize_t n = ::recv(fd, tmp, sizeof tmp, 0);
    if (n <= 0) return false;
    buf.append(tmp, static_cast<size_t>(n));
  }
  return true;
}

// Next chunk of a chunked body; "" + true = terminal chunk.
bool read_chunk(int fd, std::string& buf, std::string* chunk, int timeout_ms) {
  size_t eol;
  while ((eol = buf.find("\r\n")) == std::string::npos)
    if (!fill(fd, buf, buf.size() + 1, timeout_ms)) return false;
  const size_t len = std::strtoul(buf.c_str(), nullptr, 16);
  buf.erase(0, eol + 2);
  if (!fill(fd, buf, len + 2, timeout_ms)) return false;
  chunk->assign(buf, 0, len);
  buf.erase(0, len + 2);
  return true;
}

}  // namespace

// ================================================================ connection
EtcdCoord::~EtcdCoord() { close(); }

int EtcdCoord::dial(int timeout_ms) {
  for (const auto& [h, p] : endpoints_) {
    std::string err;
    const int fd = net::tcp_connect(h, p, timeout_ms, &err);
    if (fd >= 0) return fd;
  }
  return -1;
}

ErrorCode EtcdCoord::connect(const std::string& endpoints, int timeout_ms) {
  std::stringstream ss(endpoints);
  std::string ep;
  endpoints_.clear();
  while (std::getline(ss, ep, ',')) {
    for (const char* pfx : {"etcd://", "http://"})
      if (ep.compare(0, std::strlen(pfx), pfx) == 0) ep = ep.substr(std::strlen(pfx));
    while (!ep.empty() && (ep.back() == '/' || ep.back() == ' ')) ep.pop_back();
    auto hp = split_host_port(ep);
    if (hp) endpoints_.emplace_back(hp->first, static_cast<uint16_t>(hp->second));
  }
  closing_.store(false);
  {
    std::lock_guard<std::mutex> lk(mu_);
    if (fd_ >= 0) ::close(fd_);
    fd_ = dial(timeout_ms);
    inbuf_.clear();
    if (fd_ < 0) return ErrorCode::CONNECTION_FAILED;
  }
  // a real round trip: is this an etcd v3 gateway?
  Json q = Json::object();
  q["key"] = b64_encode("/");
  auto r = post("/v3/kv/range", q);
  if (!r.ok()) return ErrorCode::ETCD_ERROR;
  return ErrorCode::OK;
}

Result<Json> EtcdCoord::post(const std::string& path, const Json& body) {
  const std::string payload = body.dump();
  std::lock_guard<std::mutex> lk(mu_);
  requests_.fetch_add(1);
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (fd_ < 0) {
      fd_ = dial(1000);
      inbuf_.clear();
      if (fd_ < 0) return ErrorCode::ETCD_ERROR;
    }
    const std::string req = "POST " + path + " HTTP/1.1\r\nHost: " + endpoints_.front().first + "\r\nContent-Type: application/json\r\nContent-Length: " +
                            std::to_string(payload.size()) + "\r\nConnection: keep-alive\r\n\r\n" + payload;
    int status = 0;
    bool chunked = false;
    int64_t clen = -1;
    std::string resp;
    bool ok = net::send_all(fd_, req.data(), req.size(), 5000) && read_head(fd_, inbuf_, &status, &chunked, &clen, 10000);
    if (ok) {
      if (chunked) {
        std::string c;
        do {
          ok = read_chunk(fd_, inbuf_, &c, 10000);
          resp += c;
        } while (ok && !c.empty());
      } else if (clen >= 0) {
        ok = fill(fd_, inbuf_, static_cast<size_t>(clen), 10000);
        if (ok) {
          resp.assign(inbuf_, 0, static_cast<size_t>(clen));
          inbuf_.erase(0, static_cast<size_t>(clen));
        }
      } else {
        ok = false;  // neither framing: cannot keep the connection in sync
      }
    }
    if (!ok) {
      ::close(fd_);
      fd_ = -1;
      if (closing_.load()) return ErrorCode::ETCD_ERROR;
      continue;  // the server closed an idle keep-alive connection (or went away): one retry on a fresh one
    }
    std::string perr;
    auto j = Json::parse(resp, &perr);
    if (!j) return ErrorCode::ETCD_ERROR;
    if (status != 200) {
      BB_VLOG(1) << "etcd " << path << " -> HTTP " << status << ": " << resp.substr(0, 200);
      // grpc-gateway errors: {"error": "...", "code": N}; code 5 = NOT_FOUND (e.g. an unknown lease)
      const int64_t code = num(j->at("code"));
      return code == 5 ? ErrorCode::ETCD_LEASE_ERROR : ErrorCode::ETCD_ERROR;
    }
    return *j;
  }
  return ErrorCode::ETCD_ERROR;
}

void EtcdCoord::close() {
  closing_.store(true);
  std::vector<std::shared_ptr<Watch>> ws;
  {
    std::lock_guard<std::mutex> lk(wmu_);
    for (auto& [id, w] : watches_) ws.push_back(w);
    watches_.clear();
  }
  for (auto& w : ws) unwatch_impl(w);
  std::lock_guard<std::mutex> lk(mu_);
  if (fd_ >= 0) ::close(fd_);
  fd_ = -1;
}

// ================================================================ KV
ErrorCode EtcdCoord::put(const std::string& key, const std::string& value, LeaseId lease) {
  if (key.empty()) return ErrorCode::INVALID_KEY;
  auto r = post("/v3/kv/put", put_op(key, value, lease));
  return r.ok() ? ErrorCode::OK : r.error();
}

Result<KeyValue> EtcdCoord::get_kv(const std::string& key) {
  Json q = Json::object();
  q["key"] = b64_encode(key);
  auto r = post("/v3/kv/range", q);
  if (!r.ok()) return r.error();
  const Json& kvs = r.value().at("kvs");
  if (!kvs.is_array() || kvs.size() == 0) return ErrorCode::ETCD_KEY_NOT_FOUND;
  return kv_from(kvs.as_array()[0]);
}

ErrorCode EtcdCoord::del(const std::string& key) {
  Json q = Json::object();
  q["key"] = b64_encode(key);
  auto r = post("/v3/kv/deleterange", q);
  return r.ok() ? ErrorCode::OK : r.error();
}

Result<std::vector<KeyValue>> EtcdCoord::get_with_prefix(const std::string& prefix) {
  Json q = Json::object();
  q["key"] = b64_encode(prefix.empty() ? std::string(1, '\0') : prefix);
  q["range_end"] = b64_encode(prefix_end(prefix));
  auto r = post("/v3/kv/range", q);
  if (!r.ok()) return r.error();
  std::vector<KeyValue> out;
  const Json& kvs = r.value().at("kvs");
  if (kvs.is_array())
    for (const auto& j : kvs.as_array()) out.push_back(kv_from(j));
  return out;
}

Result<size_t> EtcdCoord::del_prefix(const std::string& prefix) {
  Json q = Json::object();
  q["key"] = b64_encode(prefix.empty() ? std::string(1, '\0') : prefix);
  q["range_end"] = b64_encode(prefix_end(prefix));
  auto r = post("/v3/kv/deleterange", q);
  if (!r.ok()) return r.error();
  return static_cast<size_t>(num(r.value().at("deleted")));
}

int64_t EtcdCoord::revision() {
  Json q = Json::object();
  q["key"] = b64_encode("/");
  q["count_only"] = true;
  auto r = post("/v3/kv/range", q);
  if (!r.ok()) return -1;
  return num(r.value().at("header").at("revision"));
}

// ================================================================ leases
Result<LeaseId> EtcdCoord::grant_lease(int64_t ttl_sec) {
  if (ttl_sec <= 0) return ErrorCode::INVALID_PARAMETERS;
  Json q = Json::object();
  q["TTL"] = std::to_string(ttl_sec);
  auto r = post("/v3/lease/grant", q);
  if (!r.ok()) return r.error();
  const LeaseId id = num(r.value().at("ID"));
  if (id == 0) return ErrorCode::ETCD_LEASE_ERROR;
  return id;
}

ErrorCode EtcdCoord::keep_alive(LeaseId lease) {
  Json q = Json::object();
  q["ID"] = std::to_string(lease);
  auto r = post("/v3/lease/keepalive", q);
  if (!r.ok()) return r.error();
  // {"result": {"header":..., "ID": "...", "TTL": "n"}}; an expired / unknown lease comes back without a TTL
  const Json& res = r.value().contains("result") ? r.value().at("result") : r.value();
  return num(res.at("TTL")) > 0 ? ErrorCode::OK : ErrorCode::ETCD_LEASE_ERROR;
}

ErrorCode EtcdCoord::revoke_lease(LeaseId lease) {
  Json q = Json::object();
  q["ID"] = std::to_string(lease);
  auto r = post("/v3/lease/revoke", q);
  if (!r.ok() && r.error() == ErrorCode::ETCD_ERROR) r = post("/v3/kv/lease/revoke", q);  // etcd 3.3 path
  return r.ok() ? ErrorCode::OK : ErrorCode::ETCD_LEASE_ERROR;
}

Result<int64_t> EtcdCoord::lease_remaining_ms(LeaseId lease) {
  Json q = Json::object();
  q["ID"] = std::to_string(lease);
  auto r = post("/v3/lease/timetolive", q);
  if (!r.ok()) return r.error();
  const int64_t ttl = num(r.value().at("TTL"));
  if (ttl < 0) return ErrorCode::ETCD_LEASE_ERROR;  // -1: expired
  return ttl * 1000;
}

// ================================================================ transactions
Result<bool> EtcdCoord::txn(const Json& compare, const Json& success_op) {
  Json q = Json::object();
  Json cmp = Json::array();
  cmp.push_back(compare);
  Json ok = Json::array();
  ok.push_back(success_op);
  q["compare"] = cmp;
  q["success"] = ok;
  auto r = post("/v3/kv/txn", q);
  if (!r.ok()) return r.error();
  return r.value().at("succeeded").as_bool(false);
}

namespace {
Json cmp_create(const std::string& key, int64_t rev) {
  Json c = Json::object();
  c["key"] = b64_encode(key);
  c["target"] = "CREATE";
  c["result"] = "EQUAL";
  c["create_revision"] = std::to_string(rev);
  return c;
}
Json cmp_value(const std::string& key, const std::string& value) {
  Json c = Json::object();
  c["key"] = b64_encode(key);
  c["target"] = "VALUE";
  c["result"] = "EQUAL";
  c["value"] = b64_encode(value);
  return c;
}
Json op_put(const std::string& key, const std::string& value, LeaseId lease) {
  Json o = Json::object();
  o["request_put"] = put_op(key, value, lease);
  return o;
}
Json op_del(const std::string& key) {
  Json d = Json::object();
  d["key"] = b64_encode(key);
  Json o = Json::object();
  o["request_delete_range"] = d;
  return o;
}
}  // namespace

Result<bool> EtcdCoord::put_if_absent(const std::string& key, const std::string& value, LeaseId lease) {
  return txn(cmp_create(key, 0), op_put(key, value, lease));
}
Result<bool> EtcdCoord::compare_and_swap(const std::string& key, const std::string& expected, const std::string& value, LeaseId lease) {
  return txn(cmp_value(key, expected), op_put(key, value, lease));
}
Result<bool> EtcdCoord::compare_and_delete(const std::string& key, const std::string& expected) {
  return txn(cmp_value(key, expected), op_del(key));
}
Result<bool> EtcdCoord::guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                                    const std::string& value) {
  return txn(cmp_create(guard_key, guard_create_revision), op_put(key, value, 0));
}
Result<bool> EtcdCoord::guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) {
  return txn(cmp_create(guard_key, guard_create_revision), op_del(key));
}

// ================================================================ watches
struct EtcdCoord::Watch {
  std::string prefix;
  WatchCallback cb;
  std::atomic<int> fd{-1};
  std::atomic<bool> stop{false};
  std::thread thread;
  std::mutex mu;
  std::condition_variable cv;
  bool created = false, failed = false;
};

void EtcdCoord::watch_loop(std::shared_ptr<Watch> w) {
  // One connection per watch; when it breaks (etcd restart, idle proxy) it is re-created and the prefix re-listed is NOT
  // needed: etcd replays from `start_revision`, which we advance with every event seen.
  int64_t next_rev = 0;
  while (!w->stop.load() && !closing_.load()) {
    const int fd = dial(1000);
    if (fd < 0) {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        if (!w->created) w->failed = true;
        w->cv.notify_all();
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(200));
      continue;
    }
    w->fd.store(fd);
    Json cr = Json::object();
    cr["key"] = b64_encode(w->prefix.empty() ? std::string(1, '\0') : w->prefix);
    cr["range_end"] = b64_encode(prefix_end(w->prefix));
    cr["prev_kv"] = true;
    if (next_rev > 0) cr["start_revision"] = std::to_string(next_rev);
    Json q = Json::object();
    q["create_request"] = cr;
    const std::string payload = q.dump();
    const std::string req = "POST /v3/watch HTTP/1.1\r\nHost: " + endpoints_.front().first + "\r\nContent-Type: application/json\r\nContent-Length: " +
                            std::to_string(payload.size()) + "\r\n\r\n" + payload;
    std::string buf;
    int status = 0;
    bool chunked = false;
    int64_t clen = -1;
    bool ok = net::send_all(fd, req.data(), req.size(), 5000) && read_head(fd, buf, &status, &chunked, &clen, 10000) && status == 200 && chunked;
    std::string pending;  // a message may span chunks: split on newlines / balanced JSON
    while (ok && !w->stop.load()) {
      std::string chunk;
      // poll with a short timeout so that stop is noticed; read_chunk blocks up to its timeout
      pollfd pf{fd, POLLIN, 0};
      if (buf.empty() && ::poll(&pf, 1, 200) == 0) continue;
      if (!read_chunk(fd, buf, &chunk, 10000) || chunk.empty()) break;
      pending += chunk;
      size_t nl;
      while ((nl = pending.find('\n')) != std::string::npos || (!pending.empty() && Json::parse(pending))) {
        const std::string msg = nl != std::string::npos ? pending.substr(0, nl) : pending;
        pending.erase(0, nl != std::string::npos ? nl + 1 : pending.size());
        if (msg.find_first_not_of(" \r\t") == std::string::npos) continue;
        auto j = Json::parse(msg);
        if (!j) continue;
        const Json& res = j->contains("result") ? j->at("result") : *j;
        if (res.at("created").as_bool(false)) {
          std::lock_guard<std::mutex> lk(w->mu);
          w->created = true;
          w->cv.notify_all();
        }
        const Json& evs = res.at("events");
        if (!evs.is_array()) continue;
        for (const auto& e : evs.as_array()) {
          WatchEvent ev;
          ev.type = e.at("type").as_string() == "DELETE" ? EventType::DELETE : EventType::PUT;
          const KeyValue kv = kv_from(e.at("kv"));
          ev.key = kv.key;
          ev.revision = kv.mod_revision;
          if (ev.type == EventType::DELETE) {
            if (e.contains("prev_kv")) ev.value = kv_from(e.at("prev_kv")).value;  // last value, like MemCoord
          } else {
            ev.value = kv.value;
          }
          next_rev = std::max(next_rev, kv.mod_revision + 1);
          if (!w->stop.load()) {
            try {
              w->cb(ev);
            } catch (const std::exception& ex) {
              BB_LOG(ERROR) << "watch callback threw: " << ex.what();
            }
          }
        }
      }
    }
    w->fd.store(-1);
    ::close(fd);
    {
      std::lock_guard<std::mutex> lk(w->mu);
      if (!w->created) w->failed = true;
      w->cv.notify_all();
    }
    if (!w->stop.load() && !closing_.load()) std::this_thread::sleep_for(std::chrono::milliseconds(100));
  }
}

Result<int64_t> EtcdCoord::watch_prefix(const std::string& prefix, WatchCallback cb) {
  if (!cb) return ErrorCode::ETCD_WATCH_ERROR;
  auto w = std::make_shared<Watch>();
  w->prefix = prefix;
  w->cb = std::move(cb);
  w->thread = std::thread([this, w] { watch_loop(w); });
  {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait_for(lk, std::chrono::seconds(5), [&] { return w->created || w->failed; });
    if (!w->created) {
      lk.unlock();
      unwatch_impl(w);
      return ErrorCode::ETCD_WATCH_ERROR;
    }
  }
  std::lock_guard<std::mutex> lk(wmu_);
  const int64_t id = next_watch_++;
  watches_[id] = w;
  return id;
}

void EtcdCoord::unwatch_impl(const std::shared_ptr<Watch>& w) {
  w->stop.store(true);
  const int fd = w->fd.load();
  if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
  if (w->thread.joinable()) {
    if (w->thread.get_id() == std::this_thread::get_id()) w->thread.detach();  // unwatch from inside the callback
    else w->thread.join();
  }
}

ErrorCode EtcdCoord::unwatch(int64_t watch_id) {
  std::shared_ptr<Watch> w;
  {
    std::lock_guard<std::mutex> lk(wmu_);
    auto it = watches_.find(watch_id);
    if (it == watches_.end()) return ErrorCode::ETCD_WATCH_ERROR;
    w = it->second;
    watches_.erase(it);
  }
  unwatch_impl(w);  // barrier: the thread has been joined, no callback is running
  return ErrorCode::OK;
}

}  // namespace bb::coord
