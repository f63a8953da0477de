// CXL transport / pool configuration (reference include/blackbird/transport/cxl_transport_config.h:13-118,
// configs/cxl_worker.yaml).  In the reference these types are declared and printed by an example but nothing
// consumes them; here the worker config loader parses them (`transport:` block and per-pool `config:` blocks),
// the CXL backend takes its interleave / NUMA / persistence settings from CxlMemoryPoolConfig, and the worker
// advertises `resolve_interconnects()` of its transport block to the placement engine.
#pragma once
#include <cstdint>
#include <optional>
#include <string>
#include <vector>

#include "common/json.h"

namespace bb {

enum class CxlInterconnectType : uint32_t { CXL_MEM = 0, CXL_CACHE = 1, CXL_IO = 2, CXL_FABRIC = 3, HYBRID = 4 };
enum class CxlTransportProtocol : uint32_t { DIRECT_CXL = 0, RDMA_OVER_CXL = 1, NVLINK = 2, CUSTOM = 999 };

const char* to_string(CxlInterconnectType t);
const char* to_string(CxlTransportProtocol p);
std::optional<CxlInterconnectType> parse_cxl_interconnect(const std::string& s);   // "cxl_mem", "CXL.mem", "cxl_fabric", ...
std::optional<CxlTransportProtocol> parse_cxl_protocol(const std::string& s);      // "direct_cxl", "rdma_over_cxl", "nvlink", "custom"

struct CxlTransportConfig {
  CxlInterconnectType interconnect_type = CxlInterconnectType::CXL_MEM;
  CxlTransportProtocol transport_protocol = CxlTransportProtocol::DIRECT_CXL;
  bool enable_fabric_manager = false;
  std::string fabric_manager_endpoint;
  std::vector<std::string> fabric_devices;
  uint64_t max_transfer_size = 1ull << 30;
  uint32_t queue_depth = 128;
  bool enable_zero_copy = true;
  bool enable_multipath = false;
  std::vector<std::string> fallback_transports;  // tried in order when the primary protocol is unavailable
  uint32_t priority = 0;
  uint32_t bandwidth_limit_gbps = 0;             // 0 = unlimited
  bool enable_cxl_hdm = true;
  bool enable_cxl_switch = false;
  std::string cxl_port_id;

  static CxlTransportConfig from_json(const Json& j);
  Json to_json() const;
  // Interconnect names this worker should advertise: the primary protocol first, then the fallbacks
  // ("nvlink" only when `have_gpu`; "cxl" only when a dax/cxl device is `present`; "tcp" always last).
  std::vector<std::string> resolve_interconnects(bool cxl_present, bool have_gpu) const;
};

struct CxlMemoryPoolConfig {
  std::string device_id;
  std::string device_path;   // /dev/cxl/memN
  std::string dax_device;    // /dev/daxX.Y (mapped when openable; anonymous placeholder otherwise)
  uint64_t capacity = 0;
  uint32_t latency_ns = 0;
  uint32_t bandwidth_gbps = 0;
  bool is_persistent = false;
  bool supports_cache_coherency = true;
  bool enable_numa_binding = false;
  int numa_node = -1;
  std::vector<int> cpu_affinity;
  uint64_t interleave_ways = 1;
  uint64_t interleave_granularity = 256;
  uint32_t cache_line_size = 64;

  static CxlMemoryPoolConfig from_json(const Json& j);
  Json to_json() const;
};

// Size-based tier preference (reference cxl_worker.yaml `allocation.preferred_tiers`, parsed by nothing there):
// objects whose WorkerConfig names no preferred class get the classes whose [min_size, max_size] window
// contains the object size, in list order.
struct TierRule {
  std::string storage_class;  // canonical class name
  uint64_t min_size = 0;
  uint64_t max_size = UINT64_MAX;
};
std::vector<TierRule> tier_rules_from_json(const Json& j);
std::vector<std::string> tier_classes_for_size(const std::vector<TierRule>& rules, uint64_t size);

}  // namespace bb
