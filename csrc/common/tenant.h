// Tenants: named principals with their own secret, key-prefix grants and an admission budget.
//
// The reference has one trust level -- whoever reaches the port may do anything -- and lists "Security (mTLS), ACLs" and
// "admission control" as roadmap / Keystone duties (README.md:104-108, 146-153).  The cluster token (net/tcp.h) made that
// "members and everyone else"; tenants split the members' clients further:
//
//   * identity   a tenant proves its own secret in the RPC handshake ("BBT1" / "BBT2" hello, net/tcp.h), so a server knows
//                WHO is calling without any member secret being handed to applications;
//   * ACL        read / write grants are key prefixes ("ckpt/llama/", "" or "*" = everything); a write grant implies read.
//                The Keystone checks every key of every object call against them (rpc/rpc_service.cpp) and a tenant
//                connection is confined to the object-level methods -- worker registration, migration, drain, scrub,
//                remove_all ... stay with the members unless the tenant is marked `admin`;
//   * admission  `quota_bytes` / `max_objects` bound what a tenant may have placed at any time (size x replicas, counted
//                from put_start to the object's removal, expiry or eviction; rebuilt from the metadata log after a
//                fail-over).  A put that would cross the line is refused with QUOTA_EXCEEDED before anything is allocated.
//
// The table lives in a YAML file every server process reads (`tenants_file:` in keystone / worker YAML, --tenants-file,
// BB_TENANTS_FILE); SIGHUP-free reload: the file's mtime is polled by the owner (KeystoneService health loop / worker
// heartbeat) through reload_tenants_if_changed().  A client names itself with BlackbirdClientOptions::tenant /
// tenant_secret, `--tenant NAME --tenant-secret S` in the tools, or BB_TENANT / BB_TENANT_SECRET.
//
//   tenants:
//     - name: trainer
//       secret: "s3cr3t"            # or secret_env: TRAINER_SECRET
//       write: ["ckpt/"]
//       read:  ["datasets/", "ckpt/"]
//       quota_bytes: 512GB
//       max_objects: 100000
//     - name: ops
//       secret_env: OPS_SECRET
//       admin: true
//       write: ["*"]
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "common/error.h"

namespace bb {

struct Tenant {
  std::string name;
  std::string secret;
  std::vector<std::string> read_prefixes;
  std::vector<std::string> write_prefixes;
  uint64_t quota_bytes = 0;  // 0 = unlimited
  uint64_t max_objects = 0;  // 0 = unlimited
  bool admin = false;        // may call the cluster-management methods as well

  bool may_write(std::string_view key) const;
  bool may_read(std::string_view key) const;  // read grants and write grants
  // A listing of `prefix` shows only keys the tenant may read iff one grant covers the whole prefix.
  bool may_list(std::string_view prefix) const;
};

constexpr size_t kMaxTenantName = 64;

// Process-wide table (like the cluster token): servers look callers up in it, Keystone reads budgets from it.
void set_tenants(std::vector<Tenant> tenants);
std::shared_ptr<const Tenant> find_tenant(std::string_view name);
std::vector<std::string> tenant_names();
// Parses the YAML above; on error the current table is left as it was.
ErrorCode load_tenants_text(std::string_view yaml, std::string* err = nullptr);
ErrorCode load_tenants_file(const std::string& path, std::string* err = nullptr);
// Re-reads the file given to the last load_tenants_file() (or BB_TENANTS_FILE on first use) when its mtime or size
// changed.  Returns true if a new table was installed.
bool reload_tenants_if_changed();

// Client side: the identity this process presents when it holds no member token.
void set_client_tenant(const std::string& name, const std::string& secret);
std::pair<std::string, std::string> client_tenant();  // (name, secret); env BB_TENANT / BB_TENANT_SECRET by default

// The tenant on whose behalf the current thread is executing a Keystone call (set by the RPC layer around the handler;
// in-process users may set it themselves).  nullptr = a member / in-process caller: no ACL, no budget.
class TenantScope {
 public:
  explicit TenantScope(std::shared_ptr<const Tenant> t);
  ~TenantScope();
  TenantScope(const TenantScope&) = delete;
  TenantScope& operator=(const TenantScope&) = delete;

 private:
  std::shared_ptr<const Tenant> prev_;
};
const Tenant* current_tenant();

}  // namespace bb
