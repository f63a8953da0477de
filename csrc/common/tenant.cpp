#include "common/tenant.h"

#include <sys/stat.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <optional>

#include "common/json.h"
#include "common/log.h"
#include "common/yaml.h"

namespace bb {
namespace {
bool covers(const std::vector<std::string>& grants, std::string_view key) {
  for (const auto& g : grants)
    if (g.empty() || g == "*" || (key.size() >= g.size() && key.compare(0, g.size(), g) == 0)) return true;
  return false;
}

using Table = std::map<std::string, std::shared_ptr<const Tenant>, std::less<>>;
std::mutex g_mu;
std::shared_ptr<const Table> g_table = std::make_shared<Table>();
std::string g_file;  // the file behind the table (reload)
bool g_file_init = false;
int64_t g_file_mtime_ns = 0;
int64_t g_file_size = -1;

std::string g_client_name, g_client_secret;
bool g_client_init = false;

thread_local std::shared_ptr<const Tenant> t_current;

bool stat_file(const std::string& path, int64_t* mtime_ns, int64_t* size) {
  struct stat st{};
  if (::stat(path.c_str(), &st) != 0) return false;
  *mtime_ns = static_cast<int64_t>(st.st_mtim.tv_sec) * 1000000000ll + st.st_mtim.tv_nsec;
  *size = static_cast<int64_t>(st.st_size);
  return true;
}

bool read_prefix_list(const Json& j, std::vector<std::string>* out, std::string* err, const std::string& who, const char* field) {
  if (j.is_null()) return true;
  if (j.is_string()) {
    out->push_back(j.as_string());
    return true;
  }
  if (!j.is_array()) {
    if (err) *err = "tenant " + who + ": `" + field + "` must be a list of key prefixes";
    return false;
  }
  for (const auto& e : j.as_array()) {
    if (!e.is_string()) {
      if (err) *err = "tenant " + who + ": `" + field + "` entries must be strings";
      return false;
    }
    out->push_back(e.as_string());
  }
  return true;
}

std::optional<std::vector<Tenant>> parse_table(const Json& doc, std::string* err) {
  const Json& list = doc.is_object() ? doc.at("tenants") : doc;
  std::vector<Tenant> out;
  if (list.is_null()) return out;  // an empty file: no tenants
  if (!list.is_array()) {
    if (err) *err = "`tenants` must be a list";
    return std::nullopt;
  }
  for (const auto& j : list.as_array()) {
    Tenant t;
    if (!j.is_object() || !j.at("name").is_string()) {
      if (err) *err = "every tenant needs a `name`";
      return std::nullopt;
    }
    t.name = j.at("name").as_string();
    if (t.name.empty() || t.name.size() > kMaxTenantName) {
      if (err) *err = "tenant name must be 1.." + std::to_string(kMaxTenantName) + " bytes";
      return std::nullopt;
    }
    for (const auto& o : out)
      if (o.name == t.name) {
        if (err) *err = "tenant " + t.name + " is listed twice";
        return std::nullopt;
      }
    if (j.contains("secret")) t.secret = j.at("secret").as_string();
    if (t.secret.empty() && j.contains("secret_env")) {
      if (const char* e = std::getenv(j.at("secret_env").as_string().c_str())) t.secret = e;
    }
    if (t.secret.empty()) {  // an identity nobody can prove would be a tenant anybody can claim once a bug creeps in: refuse
      if (err) *err = "tenant " + t.name + " has no secret (`secret:` or a set `secret_env:`)";
      return std::nullopt;
    }
    if (!read_prefix_list(j.at("read"), &t.read_prefixes, err, t.name, "read")) return std::nullopt;
    if (!read_prefix_list(j.at("write"), &t.write_prefixes, err, t.name, "write")) return std::nullopt;
    if (j.contains("quota_bytes")) {
      const Json& q = j.at("quota_bytes");
      std::optional<uint64_t> v = q.is_string() ? parse_size(q.as_string()) : std::optional<uint64_t>(static_cast<uint64_t>(std::max<int64_t>(0, q.as_int(0))));
      if (!v) {
        if (err) *err = "tenant " + t.name + ": quota_bytes is not a size";
        return std::nullopt;
      }
      t.quota_bytes = *v == UINT64_MAX ? 0 : *v;  // "unlimited"
    }
    if (j.contains("max_objects")) t.max_objects = static_cast<uint64_t>(std::max<int64_t>(0, j.at("max_objects").as_int(0)));
    t.admin = j.at("admin").as_bool(false);
    out.push_back(std::move(t));
  }
  return out;
}
}  // namespace

bool Tenant::may_write(std::string_view key) const { return covers(write_prefixes, key); }
bool Tenant::may_read(std::string_view key) const { return covers(read_prefixes, key) || covers(write_prefixes, key); }
bool Tenant::may_list(std::string_view prefix) const { return may_read(prefix); }  // a grant that is a prefix of `prefix` covers all of it

void set_tenants(std::vector<Tenant> tenants) {
  auto t = std::make_shared<Table>();
  for (auto& x : tenants) {
    std::string name = x.name;
    (*t)[name] = std::make_shared<const Tenant>(std::move(x));
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_table = std::move(t);
}

std::shared_ptr<const Tenant> find_tenant(std::string_view name) {
  std::shared_ptr<const Table> t;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    t = g_table;
  }
  auto it = t->find(name);
  return it == t->end() ? nullptr : it->second;
}

std::vector<std::string> tenant_names() {
  std::shared_ptr<const Table> t;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    t = g_table;
  }
  std::vector<std::string> v;
  for (const auto& [n, _] : *t) v.push_back(n);
  return v;
}

ErrorCode load_tenants_text(std::string_view yaml, std::string* err) {
  auto doc = parse_yaml(yaml, err);
  if (!doc) return ErrorCode::INVALID_CONFIGURATION;
  auto parsed = parse_table(*doc, err);
  if (!parsed) return ErrorCode::INVALID_CONFIGURATION;
  set_tenants(std::move(*parsed));
  return ErrorCode::OK;
}

ErrorCode load_tenants_file(const std::string& path, std::string* err) {
  std::string e;
  auto doc = load_yaml_file(path, &e);
  int64_t mt = 0, sz = -1;
  stat_file(path, &mt, &sz);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_file = path;
    g_file_init = true;
    g_file_mtime_ns = mt;  // a file that does not parse is not re-read until it changes again
    g_file_size = sz;
  }
  if (!doc) {
    if (err) *err = path + ": " + e;
    return ErrorCode::INVALID_CONFIGURATION;
  }
  auto parsed = parse_table(*doc, &e);
  if (!parsed) {
    if (err) *err = path + ": " + e;
    return ErrorCode::INVALID_CONFIGURATION;
  }
  const size_t n = parsed->size();
  set_tenants(std::move(*parsed));
  BB_LOG(INFO) << "tenants: " << n << " loaded from " << path;
  return ErrorCode::OK;
}

bool reload_tenants_if_changed() {
  std::string path;
  int64_t old_mt, old_sz;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_file_init) {
      g_file_init = true;
      if (const char* e = std::getenv("BB_TENANTS_FILE")) g_file = e;
      g_file_size = -1;
    }
    path = g_file;
    old_mt = g_file_mtime_ns;
    old_sz = g_file_size;
  }
  if (path.empty()) return false;
  int64_t mt = 0, sz = -1;
  if (!stat_file(path, &mt, &sz)) return false;  // gone: keep what we have (an editor's rename window must not open the cluster)
  if (mt == old_mt && sz == old_sz) return false;
  std::string err;
  if (load_tenants_file(path, &err) != ErrorCode::OK) {
    BB_LOG(ERROR) << "tenants: " << err << " -- keeping the previous table";
    return false;
  }
  return true;
}

void set_client_tenant(const std::string& name, const std::string& secret) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_client_name = name;
  g_client_secret = secret;
  g_client_init = true;
}

std::pair<std::string, std::string> client_tenant() {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_client_init) {
    if (const char* e = std::getenv("BB_TENANT")) g_client_name = e;
    if (const char* e = std::getenv("BB_TENANT_SECRET")) g_client_secret = e;
    g_client_init = true;
  }
  return {g_client_name, g_client_secret};
}

TenantScope::TenantScope(std::shared_ptr<const Tenant> t) : prev_(std::move(t_current)) { t_current = std::move(t); }
TenantScope::~TenantScope() { t_current = std::move(prev_); }
const Tenant* current_tenant() { return t_current.get(); }

}  // namespace bb
