#include "common/types.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <random>
#include <stdexcept>

#include "common/yaml.h"

namespace bb {

std::string_view to_string(StorageClass c) noexcept {
  switch (c) {
    case StorageClass::STORAGE_UNSPECIFIED: return "STORAGE_UNSPECIFIED";
    case StorageClass::RAM_CPU: return "RAM_CPU";
    case StorageClass::RAM_GPU: return "RAM_GPU";
    case StorageClass::NVME: return "NVME";
    case StorageClass::SSD: return "SSD";
    case StorageClass::HDD: return "HDD";
    case StorageClass::CXL_MEMORY: return "CXL_MEMORY";
    case StorageClass::CXL_TYPE2_DEVICE: return "CXL_TYPE2_DEVICE";
    case StorageClass::CUSTOM: return "CUSTOM";
  }
  return "UNKNOWN";
}

std::optional<StorageClass> parse_storage_class(std::string_view sv) noexcept {
  std::string s;
  for (char c : sv) s += static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
  if (s == "RAM_CPU" || s == "DRAM" || s == "RAM") return StorageClass::RAM_CPU;
  if (s == "RAM_GPU" || s == "GPU" || s == "HBM") return StorageClass::RAM_GPU;
  if (s == "NVME") return StorageClass::NVME;
  if (s == "SSD") return StorageClass::SSD;
  if (s == "HDD") return StorageClass::HDD;
  if (s == "CXL_MEMORY" || s == "CXL") return StorageClass::CXL_MEMORY;
  if (s == "CXL_TYPE2_DEVICE" || s == "CXL_TYPE2") return StorageClass::CXL_TYPE2_DEVICE;
  if (s == "CUSTOM") return StorageClass::CUSTOM;
  if (s == "STORAGE_UNSPECIFIED" || s == "UNSPECIFIED") return StorageClass::STORAGE_UNSPECIFIED;
  // numeric form
  if (!s.empty() && std::all_of(s.begin(), s.end(), [](char c) { return std::isdigit(static_cast<unsigned char>(c)); })) {
    int v = std::atoi(s.c_str());
    switch (v) {
      case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: case 999: return static_cast<StorageClass>(v);
      default: break;
    }
  }
  return std::nullopt;
}

int tier_rank(StorageClass c) noexcept {
  switch (c) {
    case StorageClass::RAM_GPU: return 0;
    case StorageClass::RAM_CPU: return 1;
    case StorageClass::CXL_MEMORY: return 2;
    case StorageClass::CXL_TYPE2_DEVICE: return 2;
    case StorageClass::NVME: return 3;
    case StorageClass::SSD: return 4;
    case StorageClass::HDD: return 5;
    default: return 6;
  }
}

bool is_disk_class(StorageClass c) noexcept {
  return c == StorageClass::NVME || c == StorageClass::SSD || c == StorageClass::HDD;
}

// ---------------------------------------------------------------- WorkerConfig
Json to_json(const WorkerConfig& c) {
  Json j = Json::object();
  j["replication_factor"] = c.replication_factor;
  j["max_workers_per_copy"] = c.max_workers_per_copy;
  j["enable_soft_pin"] = c.enable_soft_pin;
  j["preferred_node"] = c.preferred_node;
  Json pc = Json::array();
  for (auto sc : c.preferred_classes) pc.push_back(std::string(to_string(sc)));
  j["preferred_classes"] = pc;
  j["ttl_ms"] = c.ttl_ms;
  j["enable_locality_awareness"] = c.enable_locality_awareness;
  j["prefer_contiguous"] = c.prefer_contiguous;
  j["min_shard_size"] = c.min_shard_size;
  j["checksum"] = std::string(to_string(c.checksum));
  j["pack_fp8"] = c.pack_fp8;
  j["symmetric_replicas"] = c.symmetric_replicas;
  return j;
}

WorkerConfig worker_config_from_json(const Json& j) {
  WorkerConfig c;
  if (j.contains("replication_factor")) c.replication_factor = j.at("replication_factor").as_uint();
  if (j.contains("max_workers_per_copy")) c.max_workers_per_copy = j.at("max_workers_per_copy").as_uint();
  if (j.contains("enable_soft_pin")) c.enable_soft_pin = j.at("enable_soft_pin").as_bool();
  if (j.contains("preferred_node")) c.preferred_node = j.at("preferred_node").as_string();
  for (const auto& e : j.at("preferred_classes").as_array())
    if (auto sc = parse_storage_class(e.as_string())) c.preferred_classes.push_back(*sc);
  if (j.contains("ttl_ms")) c.ttl_ms = j.at("ttl_ms").as_uint();
  if (j.contains("enable_locality_awareness")) c.enable_locality_awareness = j.at("enable_locality_awareness").as_bool();
  if (j.contains("prefer_contiguous")) c.prefer_contiguous = j.at("prefer_contiguous").as_bool();
  if (j.contains("min_shard_size")) c.min_shard_size = j.at("min_shard_size").as_uint();
  if (j.contains("checksum")) {
    const std::string s = j.at("checksum").as_string();
    c.checksum = s == "crc32c" ? ChecksumAlgo::CRC32C : s == "none" ? ChecksumAlgo::NONE : (s == "xxh3" || s == "xxhash") ? ChecksumAlgo::XXH3 : ChecksumAlgo::BBH64;
  }
  if (j.contains("pack_fp8")) c.pack_fp8 = j.at("pack_fp8").as_bool();
  if (j.contains("symmetric_replicas")) c.symmetric_replicas = j.at("symmetric_replicas").as_bool();
  return c;
}

Json to_json(const ClusterStats& s) {
  Json j = Json::object();
  j["total_workers"] = s.total_workers;
  j["total_memory_pools"] = s.total_memory_pools;
  j["total_objects"] = s.total_objects;
  j["total_capacity"] = s.total_capacity;
  j["used_capacity"] = s.used_capacity;
  j["avg_utilization"] = s.avg_utilization;
  j["pending_objects"] = s.pending_objects;
  j["active_clients"] = s.active_clients;
  return j;
}

// ---------------------------------------------------------------- KeystoneConfig
Result<KeystoneConfig> KeystoneConfig::from_json(const Json& root, std::string* err) {
  auto fail = [&](ErrorCode ec, const std::string& m) -> Result<KeystoneConfig> {
    if (err) *err = m;
    return ec;
  };
  if (!root.is_object() || !root.contains("keystone") || !root.at("keystone").is_object())
    return fail(ErrorCode::MISSING_REQUIRED_FIELD, "missing top-level 'keystone' section");
  const Json& k = root.at("keystone");
  KeystoneConfig c;
  if (k.contains("cluster_id")) c.cluster_id = k.at("cluster_id").as_string();
  if (k.contains("service_id")) c.service_id = k.at("service_id").as_string();
  // etcd_endpoints: list or scalar (reference types.cpp:34-47)
  const Json& ee = k.contains("coord_endpoints") ? k.at("coord_endpoints") : k.at("etcd_endpoints");
  if (ee.is_array()) {
    std::string joined;
    for (const auto& e : ee.as_array()) {
      if (!joined.empty()) joined += ',';
      joined += e.as_string();
    }
    c.etcd_endpoints = joined;
  } else if (!ee.is_null()) {
    c.etcd_endpoints = ee.as_string();
  }
  if (k.contains("listen_address")) c.listen_address = k.at("listen_address").as_string();
  if (k.contains("http_metrics_port")) c.http_metrics_port = k.at("http_metrics_port").as_string();
  if (k.contains("enable_gc")) c.enable_gc = k.at("enable_gc").as_bool(true);
  if (k.contains("enable_ha")) c.enable_ha = k.at("enable_ha").as_bool(false);
  if (k.contains("eviction_ratio")) c.eviction_ratio = k.at("eviction_ratio").as_double(c.eviction_ratio);
  if (k.contains("high_watermark")) c.high_watermark = k.at("high_watermark").as_double(c.high_watermark);
  if (k.contains("client_ttl_sec")) c.client_ttl_sec = k.at("client_ttl_sec").as_int(c.client_ttl_sec);
  if (k.contains("auth_token")) c.auth_token = k.at("auth_token").as_string();
  if (k.contains("encrypt_transport")) c.encrypt_transport = k.at("encrypt_transport").as_bool();
  if (k.contains("auth_token_ro")) c.auth_token_ro = k.at("auth_token_ro").as_string();
  if (k.contains("tenants_file")) c.tenants_file = k.at("tenants_file").as_string();
  if (k.contains("http_auth_token")) c.http_auth_token = k.at("http_auth_token").as_string();
  if (k.contains("audit_log")) c.audit_log = k.at("audit_log").as_string();
  if (k.contains("worker_heartbeat_ttl_sec")) c.worker_heartbeat_ttl_sec = k.at("worker_heartbeat_ttl_sec").as_int(c.worker_heartbeat_ttl_sec);
  if (k.contains("service_registration_ttl_sec")) c.service_registration_ttl_sec = k.at("service_registration_ttl_sec").as_int(c.service_registration_ttl_sec);
  if (k.contains("service_refresh_interval_sec")) c.service_refresh_interval_sec = k.at("service_refresh_interval_sec").as_int(c.service_refresh_interval_sec);
  if (k.contains("scrub_objects_per_round")) c.scrub_objects_per_round = static_cast<int32_t>(k.at("scrub_objects_per_round").as_int(c.scrub_objects_per_round));
  if (k.contains("gc_interval_sec")) c.gc_interval_sec = k.at("gc_interval_sec").as_int(c.gc_interval_sec);
  if (k.contains("health_check_interval_sec")) c.health_check_interval_sec = k.at("health_check_interval_sec").as_int(c.health_check_interval_sec);
  if (k.contains("compaction_fragmentation_threshold")) c.compaction_fragmentation_threshold = k.at("compaction_fragmentation_threshold").as_double(0.0);
  if (k.contains("promote_after_reads")) c.promote_after_reads = static_cast<int32_t>(k.at("promote_after_reads").as_int(0));
  if (k.contains("tier_policy")) c.tier_policy = tier_rules_from_json(k.at("tier_policy"));
  if (k.contains("max_replicas")) c.max_replicas = static_cast<int32_t>(k.at("max_replicas").as_int(c.max_replicas));
  if (k.contains("default_replicas")) c.default_replicas = static_cast<int32_t>(k.at("default_replicas").as_int(c.default_replicas));
  if (k.contains("rpc_busy_poll_us")) c.rpc_busy_poll_us = static_cast<int32_t>(k.at("rpc_busy_poll_us").as_int(0));
  if (k.contains("rpc_threads")) c.rpc_threads = static_cast<int32_t>(k.at("rpc_threads").as_int(c.rpc_threads));
  if (k.contains("wal_path")) c.wal_path = k.at("wal_path").as_string();
  if (k.contains("enable_reservations")) c.enable_reservations = k.at("enable_reservations").as_bool();
  if (k.contains("reservation_ttl_ms")) c.reservation_ttl_ms = k.at("reservation_ttl_ms").as_int(c.reservation_ttl_ms);
  if (k.contains("wal_fsync")) c.wal_fsync = k.at("wal_fsync").as_bool();
  if (k.contains("wal_snapshot_mb")) c.wal_snapshot_mb = static_cast<int>(k.at("wal_snapshot_mb").as_int());
  const Json& lg = root.at("logging");
  if (lg.is_object()) {
    c.log_level = lg.at("level").as_string();
    if (lg.at("log_to_file").as_bool(false)) c.log_file = lg.at("log_file_path").as_string();
  }
  std::string verr;
  ErrorCode ec = c.validate(&verr);
  if (ec != ErrorCode::OK) return fail(ec, verr);
  return c;
}

ErrorCode KeystoneConfig::validate(std::string* err) const {
  auto fail = [&](ErrorCode ec, const char* m) {
    if (err) *err = m;
    return ec;
  };
  if (cluster_id.empty()) return fail(ErrorCode::MISSING_REQUIRED_FIELD, "keystone.cluster_id is required");
  if (listen_address.empty() || !split_host_port(listen_address)) return fail(ErrorCode::INVALID_CONFIGURATION, "keystone.listen_address must be host:port");
  if (!(eviction_ratio > 0.0 && eviction_ratio <= 1.0)) return fail(ErrorCode::VALUE_OUT_OF_RANGE, "keystone.eviction_ratio must be in (0,1]");
  if (!(high_watermark > 0.0 && high_watermark <= 1.0)) return fail(ErrorCode::VALUE_OUT_OF_RANGE, "keystone.high_watermark must be in (0,1]");
  if (gc_interval_sec <= 0 || health_check_interval_sec <= 0) return fail(ErrorCode::VALUE_OUT_OF_RANGE, "keystone intervals must be positive");
  if (worker_heartbeat_ttl_sec <= 0 || service_registration_ttl_sec <= 0 || service_refresh_interval_sec <= 0)
    return fail(ErrorCode::VALUE_OUT_OF_RANGE, "keystone TTLs must be positive");
  if (max_replicas < 1 || default_replicas < 1 || default_replicas > max_replicas)
    return fail(ErrorCode::VALUE_OUT_OF_RANGE, "keystone.default_replicas must be within [1, max_replicas]");
  return ErrorCode::OK;
}

KeystoneConfig KeystoneConfig::from_yaml(const std::string& file_path) {
  std::string err;
  auto root = load_yaml_file(file_path, &err);
  if (!root) throw std::runtime_error("Failed to load keystone config '" + file_path + "': " + err);
  auto c = from_json(*root, &err);
  if (!c.ok()) throw std::runtime_error("Invalid keystone config '" + file_path + "': " + err);
  return c.value();
}

Json to_json(const KeystoneConfig& c) {
  Json j = Json::object();
  j["cluster_id"] = c.cluster_id;
  j["etcd_endpoints"] = c.etcd_endpoints;
  j["listen_address"] = c.listen_address;
  j["http_metrics_port"] = c.http_metrics_port;
  j["service_id"] = c.service_id;
  j["enable_gc"] = c.enable_gc;
  j["enable_ha"] = c.enable_ha;
  j["eviction_ratio"] = c.eviction_ratio;
  j["high_watermark"] = c.high_watermark;
  j["client_ttl_sec"] = c.client_ttl_sec;
  j["worker_heartbeat_ttl_sec"] = c.worker_heartbeat_ttl_sec;
  j["service_registration_ttl_sec"] = c.service_registration_ttl_sec;
  j["service_refresh_interval_sec"] = c.service_refresh_interval_sec;
  j["gc_interval_sec"] = c.gc_interval_sec;
  j["scrub_objects_per_round"] = c.scrub_objects_per_round;
  j["health_check_interval_sec"] = c.health_check_interval_sec;
  j["max_replicas"] = c.max_replicas;
  j["default_replicas"] = c.default_replicas;
  return j;
}

// ---------------------------------------------------------------- MemoryPool / WorkerRecord
Json to_json(const MemoryPool& p) {
  Json j = Json::object();
  j["id"] = p.id;
  j["node_id"] = p.node_id;
  j["worker_id"] = p.worker_id;
  j["base_addr"] = p.base_addr;
  j["size"] = p.size;
  j["used"] = p.used;
  j["storage_class"] = static_cast<uint32_t>(p.storage_class);
  j["ucx_endpoint"] = p.ucx_endpoint;
  j["ucx_remote_addr"] = p.ucx_remote_addr;
  j["ucx_rkey_hex"] = p.ucx_rkey_hex;
  j["gpu_device_id"] = p.gpu_device_id;
  j["numa_node"] = p.numa_node;
  j["max_bw_gbps"] = p.max_bw_gbps;
  j["fabric_domain"] = p.fabric_domain;
  j["mount_path"] = p.mount_path;
  return j;
}

Result<MemoryPool> memory_pool_from_json(const Json& j) {
  if (!j.is_object() || !j.contains("id") || !j.contains("size")) return ErrorCode::INVALID_MEMORY_POOL;
  MemoryPool p;
  p.id = j.at("id").as_string();
  p.node_id = j.at("node_id").as_string();
  p.worker_id = j.at("worker_id").as_string();
  p.base_addr = j.at("base_addr").as_uint();
  p.size = j.at("size").as_uint();
  p.used = j.at("used").as_uint();
  const Json& sc = j.at("storage_class");
  if (sc.is_string()) {
    auto c = parse_storage_class(sc.as_string());
    if (!c) return ErrorCode::INVALID_MEMORY_POOL;
    p.storage_class = *c;
  } else {
    p.storage_class = static_cast<StorageClass>(sc.as_uint());
  }
  p.ucx_endpoint = j.at("ucx_endpoint").as_string();
  p.ucx_remote_addr = j.at("ucx_remote_addr").as_uint();
  p.ucx_rkey_hex = j.at("ucx_rkey_hex").as_string();
  p.gpu_device_id = static_cast<int32_t>(j.at("gpu_device_id").as_int(-1));
  p.numa_node = static_cast<int32_t>(j.at("numa_node").as_int(-1));
  p.max_bw_gbps = j.at("max_bw_gbps").as_double(0.0);
  p.fabric_domain = j.at("fabric_domain").as_string();
  p.mount_path = j.at("mount_path").as_string();
  if (p.id.empty() || p.size == 0) return ErrorCode::INVALID_MEMORY_POOL;
  return p;
}

Json to_json(const WorkerRecord& w) {
  Json j = Json::object();
  j["worker_id"] = w.worker_id;
  j["node_id"] = w.node_id;
  j["rpc_endpoint"] = w.rpc_endpoint;
  j["endpoint"] = w.rpc_endpoint;  // the key the reference keystone actually reads (keystone_service.cpp:772-806)
  j["ucx_endpoint"] = w.ucx_endpoint;
  Json ic = Json::array();
  for (const auto& s : w.interconnects) ic.push_back(s);
  j["interconnects"] = ic;
  Json caps = Json::object();
  Json scs = Json::array();
  for (auto sc : w.storage_classes) scs.push_back(static_cast<uint32_t>(sc));
  caps["storage_classes"] = scs;
  caps["max_bw_gbps"] = w.max_bw_gbps;
  caps["numa_node"] = w.numa_node;
  j["capabilities"] = caps;
  j["version"] = w.version;
  return j;
}

Result<WorkerRecord> worker_record_from_json(const Json& j) {
  if (!j.is_object()) return ErrorCode::INVALID_WORKER;
  WorkerRecord w;
  w.worker_id = j.at("worker_id").as_string();
  w.node_id = j.at("node_id").as_string();
  w.rpc_endpoint = j.contains("rpc_endpoint") ? j.at("rpc_endpoint").as_string() : j.at("endpoint").as_string();
  w.ucx_endpoint = j.at("ucx_endpoint").as_string();
  for (const auto& e : j.at("interconnects").as_array()) w.interconnects.push_back(e.as_string());
  const Json& caps = j.at("capabilities");
  for (const auto& e : caps.at("storage_classes").as_array()) w.storage_classes.push_back(static_cast<StorageClass>(e.as_uint()));
  w.max_bw_gbps = caps.at("max_bw_gbps").as_double(0.0);
  w.numa_node = static_cast<int32_t>(caps.at("numa_node").as_int(-1));
  w.version = j.at("version").as_string();
  return w;
}

// ---------------------------------------------------------------- helpers
std::string bytes_to_hex(const std::vector<uint8_t>& b) {
  static const char* kHex = "0123456789abcdef";
  std::string s;
  s.reserve(b.size() * 2);
  for (uint8_t v : b) {
    s += kHex[v >> 4];
    s += kHex[v & 15];
  }
  return s;
}

std::optional<std::vector<uint8_t>> hex_to_bytes(std::string_view hex) {
  if (hex.size() % 2) return std::nullopt;
  auto nib = [](char c) -> int {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
  };
  std::vector<uint8_t> out;
  out.reserve(hex.size() / 2);
  for (size_t i = 0; i < hex.size(); i += 2) {
    int a = nib(hex[i]), b = nib(hex[i + 1]);
    if (a < 0 || b < 0) return std::nullopt;
    out.push_back(static_cast<uint8_t>((a << 4) | b));
  }
  return out;
}

std::optional<std::pair<std::string, uint32_t>> split_host_port(std::string_view s) {
  size_t c = s.rfind(':');
  if (c == std::string_view::npos || c == 0 || c + 1 >= s.size()) return std::nullopt;
  std::string host(s.substr(0, c));
  std::string port(s.substr(c + 1));
  if (!std::all_of(port.begin(), port.end(), [](char ch) { return std::isdigit(static_cast<unsigned char>(ch)); })) return std::nullopt;
  unsigned long p = std::strtoul(port.c_str(), nullptr, 10);
  if (p > 65535) return std::nullopt;
  return std::make_pair(host, static_cast<uint32_t>(p));
}

UUID generate_uuid() {
  static thread_local std::mt19937_64 rng{std::random_device{}()};
  return {rng(), rng()};
}

std::string uuid_to_string(const UUID& u) {
  char buf[40];
  std::snprintf(buf, sizeof buf, "%016llx%016llx", static_cast<unsigned long long>(u.first), static_cast<unsigned long long>(u.second));
  return buf;
}

}  // namespace bb
