// Error model for blackbird_b200.
//
// Parity: reference include/blackbird/common/error/error_domain.h:14-38 (seven 1000-spaced
// domains) and error_codes.h:15-79 (code list).  Numeric values of every reference code are
// preserved so the wire protocol and logs stay comparable.  Additions: INVALID_ARGUMENT,
// ALREADY_EXISTS, NOT_FOUND (the three codes the reference's CXL backend uses but never
// declares, cxl_memory_backend.cpp:206,211,263), plus data-plane codes for the GPU fabric.
// Unlike the reference (error_codes.cpp:9-70), every code has a string and a description.
#pragma once
#include <cstdint>
#include <string_view>

namespace bb {

enum class Domain : uint32_t {
  SUCCESS = 0,
  SYSTEM = 1000,
  STORAGE = 2000,
  NETWORK = 3000,
  COORDINATION = 4000,
  DATA = 5000,
  CLIENT = 6000,
  CONFIG = 7000,
};

constexpr uint32_t domain_base(Domain d) noexcept { return static_cast<uint32_t>(d); }
constexpr bool is_in_domain(uint32_t code, Domain d) noexcept {
  return code >= domain_base(d) && code < domain_base(d) + 1000;
}

enum class ErrorCode : uint32_t {
  OK = 0,
  // SYSTEM
  INTERNAL_ERROR = 1000,
  INITIALIZATION_FAILED,
  INVALID_STATE,
  OPERATION_TIMEOUT,
  RESOURCE_EXHAUSTED,
  NOT_IMPLEMENTED,
  INVALID_ARGUMENT,
  ALREADY_EXISTS,
  NOT_FOUND,
  // STORAGE
  BUFFER_OVERFLOW = 2000,
  OUT_OF_MEMORY,
  MEMORY_POOL_NOT_FOUND,
  MEMORY_POOL_ALREADY_EXISTS,
  INVALID_MEMORY_POOL,
  ALLOCATION_FAILED,
  INSUFFICIENT_SPACE,
  MEMORY_ACCESS_ERROR,
  IO_ERROR,
  // NETWORK
  NETWORK_ERROR = 3000,
  CONNECTION_FAILED,
  TRANSFER_FAILED,
  UCX_ERROR,  // kept for numeric parity; reported for fabric (NVLink/IPC) errors
  INVALID_ADDRESS,
  REMOTE_ENDPOINT_ERROR,
  RPC_FAILED,
  FABRIC_ERROR,
  // COORDINATION
  ETCD_ERROR = 4000,
  ETCD_KEY_NOT_FOUND,
  ETCD_TRANSACTION_FAILED,
  ETCD_LEASE_ERROR,
  ETCD_WATCH_ERROR,
  LEADER_ELECTION_FAILED,
  SERVICE_REGISTRATION_FAILED,
  NOT_LEADER,
  // DATA
  OBJECT_NOT_FOUND = 5000,
  OBJECT_ALREADY_EXISTS,
  INVALID_KEY,
  INVALID_WORKER,
  WORKER_NOT_READY,
  NO_COMPLETE_WORKER,
  DATA_CORRUPTION,
  CHECKSUM_MISMATCH,
  OBJECT_NOT_READY,
  // CLIENT
  CLIENT_ERROR = 6000,
  CLIENT_NOT_FOUND,
  CLIENT_ALREADY_EXISTS,
  CLIENT_DISCONNECTED,
  SESSION_EXPIRED,
  INVALID_CLIENT_STATE,
  ACCESS_DENIED,  // the peer did not present the cluster token (net/tcp.h), or a tenant asked for a key outside its grants
  QUOTA_EXCEEDED,  // a tenant's put would exceed its budget (common/tenant.h)
  // CONFIG
  CONFIG_ERROR = 7000,
  INVALID_CONFIGURATION,
  INVALID_PARAMETERS,
  MISSING_REQUIRED_FIELD,
  VALUE_OUT_OF_RANGE,
};

constexpr Domain get_error_domain(ErrorCode c) noexcept {
  const uint32_t v = static_cast<uint32_t>(c);
  return v < 1000 ? Domain::SUCCESS : static_cast<Domain>((v / 1000) * 1000);
}
constexpr bool is_ok(ErrorCode c) noexcept { return c == ErrorCode::OK; }

std::string_view to_string(ErrorCode c) noexcept;
std::string_view get_error_description(ErrorCode c) noexcept;
std::string_view to_string(Domain d) noexcept;

}  // namespace bb
