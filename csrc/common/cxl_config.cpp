#include "common/cxl_config.h"

#include <algorithm>
#include <cctype>

#include "common/yaml.h"

namespace bb {

const char* to_string(CxlInterconnectType t) {
  switch (t) {
    case CxlInterconnectType::CXL_MEM: return "CXL.mem";
    case CxlInterconnectType::CXL_CACHE: return "CXL.cache";
    case CxlInterconnectType::CXL_IO: return "CXL.io";
    case CxlInterconnectType::CXL_FABRIC: return "CXL.fabric";
    case CxlInterconnectType::HYBRID: return "Hybrid";
  }
  return "Unknown";
}

const char* to_string(CxlTransportProtocol p) {
  switch (p) {
    case CxlTransportProtocol::DIRECT_CXL: return "Direct CXL";
    case CxlTransportProtocol::RDMA_OVER_CXL: return "RDMA over CXL";
    case CxlTransportProtocol::NVLINK: return "NVLink";
    case CxlTransportProtocol::CUSTOM: return "Custom";
  }
  return "Unknown";
}

static std::string canon(const std::string& s) {
  std::string o;
  for (char c : s)
    if (std::isalnum(static_cast<unsigned char>(c))) o.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
  return o;
}

std::optional<CxlInterconnectType> parse_cxl_interconnect(const std::string& s) {
  const std::string c = canon(s);
  if (c == "cxlmem" || c == "mem") return CxlInterconnectType::CXL_MEM;
  if (c == "cxlcache" || c == "cache") return CxlInterconnectType::CXL_CACHE;
  if (c == "cxlio" || c == "io") return CxlInterconnectType::CXL_IO;
  if (c == "cxlfabric" || c == "fabric") return CxlInterconnectType::CXL_FABRIC;
  if (c == "hybrid") return CxlInterconnectType::HYBRID;
  return std::nullopt;
}

std::optional<CxlTransportProtocol> parse_cxl_protocol(const std::string& s) {
  const std::string c = canon(s);
  if (c == "directcxl" || c == "direct") return CxlTransportProtocol::DIRECT_CXL;
  if (c == "rdmaovercxl" || c == "rdma") return CxlTransportProtocol::RDMA_OVER_CXL;
  if (c == "nvlink") return CxlTransportProtocol::NVLINK;
  if (c == "custom") return CxlTransportProtocol::CUSTOM;
  return std::nullopt;
}

static uint64_t size_of(const Json& v, uint64_t dflt) {
  if (v.is_number()) return v.as_uint();
  if (v.is_string()) {
    if (canon(v.as_string()) == "unlimited") return UINT64_MAX;
    if (auto s = parse_size(v.as_string())) return *s;
  }
  return dflt;
}

static std::vector<std::string> strings_of(const Json& v) {
  std::vector<std::string> out;
  if (v.is_array())
    for (const auto& e : v.as_array()) out.push_back(e.as_string());
  else if (v.is_string() && !v.as_string().empty())
    out.push_back(v.as_string());
  return out;
}

CxlTransportConfig CxlTransportConfig::from_json(const Json& j) {
  CxlTransportConfig c;
  if (!j.is_object()) return c;
  if (j.contains("type"))
    if (auto t = parse_cxl_interconnect(j.at("type").as_string())) c.interconnect_type = *t;
  if (j.contains("interconnect_type"))
    if (auto t = parse_cxl_interconnect(j.at("interconnect_type").as_string())) c.interconnect_type = *t;
  if (j.contains("protocol"))
    if (auto p = parse_cxl_protocol(j.at("protocol").as_string())) c.transport_protocol = *p;
  c.enable_fabric_manager = j.at("enable_fabric_manager").as_bool(c.enable_fabric_manager);
  if (j.contains("fabric_manager_endpoint")) c.fabric_manager_endpoint = j.at("fabric_manager_endpoint").as_string();
  c.fabric_devices = strings_of(j.at("fabric_devices"));
  c.max_transfer_size = size_of(j.at("max_transfer_size"), c.max_transfer_size);
  c.queue_depth = static_cast<uint32_t>(j.at("queue_depth").as_int(c.queue_depth));
  c.enable_zero_copy = j.at("enable_zero_copy").as_bool(c.enable_zero_copy);
  c.enable_multipath = j.at("enable_multipath").as_bool(c.enable_multipath);
  c.fallback_transports = strings_of(j.at("fallback_transports"));
  c.priority = static_cast<uint32_t>(j.at("priority").as_int(c.priority));
  c.bandwidth_limit_gbps = static_cast<uint32_t>(j.at("bandwidth_limit_gbps").as_int(c.bandwidth_limit_gbps));
  c.enable_cxl_hdm = j.at("enable_cxl_hdm").as_bool(c.enable_cxl_hdm);
  c.enable_cxl_switch = j.at("enable_cxl_switch").as_bool(c.enable_cxl_switch);
  if (j.contains("cxl_port_id")) c.cxl_port_id = j.at("cxl_port_id").as_string();
  return c;
}

Json CxlTransportConfig::to_json() const {
  Json j = Json::object();
  j["interconnect_type"] = to_string(interconnect_type);
  j["protocol"] = to_string(transport_protocol);
  j["enable_fabric_manager"] = enable_fabric_manager;
  j["fabric_manager_endpoint"] = fabric_manager_endpoint;
  j["max_transfer_size"] = max_transfer_size;
  j["queue_depth"] = static_cast<uint64_t>(queue_depth);
  j["enable_zero_copy"] = enable_zero_copy;
  j["enable_multipath"] = enable_multipath;
  Json fb = Json::array();
  for (const auto& f : fallback_transports) fb.push_back(f);
  j["fallback_transports"] = fb;
  j["priority"] = static_cast<uint64_t>(priority);
  j["bandwidth_limit_gbps"] = static_cast<uint64_t>(bandwidth_limit_gbps);
  j["enable_cxl_hdm"] = enable_cxl_hdm;
  j["enable_cxl_switch"] = enable_cxl_switch;
  j["cxl_port_id"] = cxl_port_id;
  return j;
}

std::vector<std::string> CxlTransportConfig::resolve_interconnects(bool cxl_present, bool have_gpu) const {
  std::vector<std::string> out;
  auto add = [&](const std::string& n) {
    if (!n.empty() && std::find(out.begin(), out.end(), n) == out.end()) out.push_back(n);
  };
  auto usable = [&](const std::string& n) {
    if (n == "nvlink") return have_gpu;
    if (n == "cxl" || n == "rdma_over_cxl") return cxl_present;
    if (n == "ucx" || n == "roce" || n == "ib") return false;  // no RDMA stack in this build: falls through to tcp
    return true;
  };
  std::string primary;
  switch (transport_protocol) {
    case CxlTransportProtocol::DIRECT_CXL: primary = "cxl"; break;
    case CxlTransportProtocol::RDMA_OVER_CXL: primary = "rdma_over_cxl"; break;
    case CxlTransportProtocol::NVLINK: primary = "nvlink"; break;
    case CxlTransportProtocol::CUSTOM: primary = "custom"; break;
  }
  if (usable(primary)) add(primary);
  for (const auto& f : fallback_transports)
    if (usable(canon(f) == "nvlink" ? "nvlink" : f)) add(canon(f) == "nvlink" ? "nvlink" : f);
  add("tcp");
  return out;
}

CxlMemoryPoolConfig CxlMemoryPoolConfig::from_json(const Json& j) {
  CxlMemoryPoolConfig c;
  if (!j.is_object()) return c;
  auto str = [&](const char* k, std::string& dst) {
    if (j.contains(k)) dst = j.at(k).as_string();
  };
  str("device_id", c.device_id);
  str("device_path", c.device_path);
  str("dax_device", c.dax_device);
  c.capacity = size_of(j.at("capacity"), 0);
  c.latency_ns = static_cast<uint32_t>(j.at("latency_ns").as_int(0));
  c.bandwidth_gbps = static_cast<uint32_t>(j.at("bandwidth_gbps").as_int(0));
  c.is_persistent = j.at("enable_persistent_mode").as_bool(j.at("is_persistent").as_bool(false));
  c.supports_cache_coherency = j.at("supports_cache_coherency").as_bool(true);
  c.enable_numa_binding = j.at("enable_numa_binding").as_bool(false);
  c.numa_node = static_cast<int>(j.at("numa_node").as_int(-1));
  if (j.at("cpu_affinity").is_array())
    for (const auto& e : j.at("cpu_affinity").as_array()) c.cpu_affinity.push_back(static_cast<int>(e.as_int(0)));
  c.interleave_ways = static_cast<uint64_t>(j.at("interleave_ways").as_int(1));
  c.interleave_granularity = static_cast<uint64_t>(j.at("interleave_granularity").as_int(256));
  c.cache_line_size = static_cast<uint32_t>(j.at("cache_line_size").as_int(64));
  return c;
}

Json CxlMemoryPoolConfig::to_json() const {
  Json j = Json::object();
  j["device_id"] = device_id;
  j["device_path"] = device_path;
  j["dax_device"] = dax_device;
  j["capacity"] = capacity;
  j["latency_ns"] = static_cast<uint64_t>(latency_ns);
  j["bandwidth_gbps"] = static_cast<uint64_t>(bandwidth_gbps);
  j["is_persistent"] = is_persistent;
  j["enable_numa_binding"] = enable_numa_binding;
  j["numa_node"] = static_cast<int64_t>(numa_node);
  j["interleave_ways"] = interleave_ways;
  j["interleave_granularity"] = interleave_granularity;
  j["cache_line_size"] = static_cast<uint64_t>(cache_line_size);
  return j;
}

std::vector<TierRule> tier_rules_from_json(const Json& j) {
  std::vector<TierRule> out;
  if (!j.is_array()) return out;
  for (const auto& e : j.as_array()) {
    if (!e.is_object() || !e.contains("storage_class")) continue;
    TierRule r;
    r.storage_class = e.at("storage_class").as_string();
    r.min_size = size_of(e.at("min_size"), 0);
    r.max_size = size_of(e.at("max_size"), UINT64_MAX);
    out.push_back(std::move(r));
  }
  return out;
}

std::vector<std::string> tier_classes_for_size(const std::vector<TierRule>& rules, uint64_t size) {
  std::vector<std::string> out;
  for (const auto& r : rules)
    if (size >= r.min_size && size <= r.max_size) out.push_back(r.storage_class);
  return out;
}

}  // namespace bb
