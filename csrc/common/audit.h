// Audit trail: who did what to the cluster, one JSON object per line, append-only.
//
// Operators of a shared store need to answer "who removed that worker" and "who keeps knocking with a wrong secret" after
// the fact; metrics count such events, the audit log names them.  Off unless a path is given -- `audit_log:` in keystone /
// worker YAML, --audit-log F, BB_AUDIT_LOG -- then every security-relevant event is appended (O_APPEND, one write per
// line, so several processes may share a file) with a wall-clock timestamp, the calling principal ("member", "ro-member"
// or the tenant's name) and the peer address:
//   auth_failed      a handshake that was refused (wrong / missing token, unknown tenant, plain client on a sealed server)
//   tenant_admitted  a connection that proved a tenant's secret
//   method_denied    a tenant or read-only member asked for a method outside its list
//   acl_denied       a tenant named a key outside its grants            (op, key)
//   quota_denied     a tenant's put would have crossed its budget        (key, bytes)
//   admin            a cluster-management call and its outcome            (op, arg, result)
//   tenants_reloaded the tenant table changed on disk and was installed   (count)
// Values are JSON-escaped; a line never exceeds a few hundred bytes (keys are cut at 256).  The reference has no
// counterpart (its roadmap lists "operability hardening", README.md:146-153).
#pragma once
#include <cstdint>
#include <initializer_list>
#include <string>
#include <string_view>
#include <utility>

namespace bb::audit {

// "" closes the log.  Returns false if the file cannot be opened for appending (the previous log stays).
bool open(const std::string& path);
bool enabled();
std::string path();
using Field = std::pair<std::string_view, std::string_view>;
// Appends {"ts": ..., "event": kind, "who": ..., "peer": ..., fields...}.  `who` / `peer` come from the calling thread's
// scope (below) unless given in `fields`.
void event(std::string_view kind, std::initializer_list<Field> fields = {});
uint64_t events_written();

// The connection on whose behalf this thread runs (set by RpcServer::dispatch / the handshake code).
class Scope {
 public:
  Scope(std::string_view who, std::string_view peer);
  ~Scope();
  Scope(const Scope&) = delete;
  Scope& operator=(const Scope&) = delete;

 private:
  std::string_view prev_who_, prev_peer_;
};

}  // namespace bb::audit
