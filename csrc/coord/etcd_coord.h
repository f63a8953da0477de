// EtcdCoord: the CoordStore interface spoken to a real etcd v3 cluster through its JSON/HTTP gateway
// (grpc-gateway: POST /v3/kv/{put,range,deleterange,txn}, /v3/lease/{grant,keepalive,revoke,timetolive}, /v3/watch).
//
// The reference talks to etcd through etcd-cpp-apiv3 + gRPC (src/etcd/etcd_service.cpp:60-86 connect, :88-204 KV and
// leases, :300-332 prefix watch; election API include/blackbird/etcd/etcd_service.h:205-232).  Neither that library nor
// gRPC exists offline, but every etcd >= 3.4 serves the same API as JSON over HTTP/1.1 on its client port, which needs
// nothing but a socket: bytes fields are base64, 64-bit integers are decimal strings, streams (watch) are chunked
// responses carrying one JSON object per message.
//
// Selected by endpoints of the form `etcd://host:port[,host:port...]` (CoordService::connect); `bb-coord` endpoints keep
// the native framed protocol.  Semantics map 1:1 onto CoordStore:
//   put_if_absent      txn { compare create_revision(key) == 0 ; success put }
//   compare_and_swap   txn { compare value(key) == expected    ; success put }         (compare_and_delete alike)
//   guarded_put / del  txn { compare create_revision(guard) == term ; success put | delete_range }   <- leader fencing
//   watch_prefix       POST /v3/watch {"create_request": {key, range_end, prev_kv}} on a dedicated connection + thread
#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common/json.h"
#include "coord/coord.h"

namespace bb::coord {

class EtcdCoord : public CoordStore {
 public:
  EtcdCoord() = default;
  ~EtcdCoord() override;
  // endpoints: comma separated "host:port" (an optional "http://" prefix is ignored); every endpoint is tried in order.
  ErrorCode connect(const std::string& endpoints, int timeout_ms = 3000);
  void close();

  ErrorCode put(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<KeyValue> get_kv(const std::string& key) override;
  ErrorCode del(const std::string& key) override;
  Result<std::vector<KeyValue>> get_with_prefix(const std::string& prefix) override;
  Result<size_t> del_prefix(const std::string& prefix) override;
  Result<LeaseId> grant_lease(int64_t ttl_sec) override;
  ErrorCode keep_alive(LeaseId lease) override;
  ErrorCode revoke_lease(LeaseId lease) override;
  Result<int64_t> lease_remaining_ms(LeaseId lease) override;
  Result<bool> put_if_absent(const std::string& key, const std::string& value, LeaseId lease = 0) override;
  Result<bool> compare_and_swap(const std::string& key, const std::string& expected, const std::string& value, LeaseId lease = 0) override;
  Result<bool> compare_and_delete(const std::string& key, const std::string& expected) override;
  Result<bool> guarded_put(const std::string& guard_key, int64_t guard_create_revision, const std::string& key,
                           const std::string& value) override;
  Result<bool> guarded_del(const std::string& guard_key, int64_t guard_create_revision, const std::string& key) override;
  Result<int64_t> watch_prefix(const std::string& prefix, WatchCallback cb) override;
  ErrorCode unwatch(int64_t watch_id) override;
  int64_t revision() override;

  uint64_t requests() const { return requests_.load(); }

 private:
  struct Watch;
  // One JSON request / response over a kept-alive connection (reconnects once when the peer went away).
  Result<Json> post(const std::string& path, const Json& body);
  Result<bool> txn(const Json& compare, const Json& success_op);
  int dial(int timeout_ms);
  void watch_loop(std::shared_ptr<Watch> w);
  void unwatch_impl(const std::shared_ptr<Watch>& w);

  std::vector<std::pair<std::string, uint16_t>> endpoints_;
  std::mutex mu_;  // serialises the request connection
  int fd_ = -1;
  std::string inbuf_;
  std::atomic<uint64_t> requests_{0};
  std::mutex wmu_;
  std::map<int64_t, std::shared_ptr<Watch>> watches_;
  int64_t next_watch_ = 1;
  std::atomic<bool> closing_{false};
};

// base64 (standard alphabet, padded) -- etcd's JSON encoding of `bytes` fields
std::string b64_encode(std::string_view raw);
bool b64_decode(std::string_view text, std::string* out);

}  // namespace bb::coord
