#include "common/error.h"

namespace bb {

namespace {
struct Entry {
  ErrorCode code;
  std::string_view name;
  std::string_view desc;
};

// One table drives both to_string and get_error_description so that no code can be
// missing from one of them (the reference forgot NOT_IMPLEMENTED / INSUFFICIENT_SPACE).
constexpr Entry kTable[] = {
    {ErrorCode::OK, "OK", "Operation completed successfully"},
    {ErrorCode::INTERNAL_ERROR, "INTERNAL_ERROR", "Unexpected internal failure"},
    {ErrorCode::INITIALIZATION_FAILED, "INITIALIZATION_FAILED", "A component failed to initialise"},
    {ErrorCode::INVALID_STATE, "INVALID_STATE", "Operation not valid in the current state"},
    {ErrorCode::OPERATION_TIMEOUT, "OPERATION_TIMEOUT", "Operation or reservation timed out"},
    {ErrorCode::RESOURCE_EXHAUSTED, "RESOURCE_EXHAUSTED", "A bounded resource is exhausted"},
    {ErrorCode::NOT_IMPLEMENTED, "NOT_IMPLEMENTED", "Feature is not implemented"},
    {ErrorCode::INVALID_ARGUMENT, "INVALID_ARGUMENT", "An argument is invalid"},
    {ErrorCode::ALREADY_EXISTS, "ALREADY_EXISTS", "Entity already exists"},
    {ErrorCode::NOT_FOUND, "NOT_FOUND", "Entity was not found"},
    {ErrorCode::BUFFER_OVERFLOW, "BUFFER_OVERFLOW", "Write exceeds the buffer"},
    {ErrorCode::OUT_OF_MEMORY, "OUT_OF_MEMORY", "Memory allocation failed"},
    {ErrorCode::MEMORY_POOL_NOT_FOUND, "MEMORY_POOL_NOT_FOUND", "Referenced memory pool is unknown"},
    {ErrorCode::MEMORY_POOL_ALREADY_EXISTS, "MEMORY_POOL_ALREADY_EXISTS", "Memory pool id already registered"},
    {ErrorCode::INVALID_MEMORY_POOL, "INVALID_MEMORY_POOL", "Memory pool descriptor is malformed"},
    {ErrorCode::ALLOCATION_FAILED, "ALLOCATION_FAILED", "Placement or range allocation failed"},
    {ErrorCode::INSUFFICIENT_SPACE, "INSUFFICIENT_SPACE", "Not enough free space in eligible pools"},
    {ErrorCode::MEMORY_ACCESS_ERROR, "MEMORY_ACCESS_ERROR", "Access outside a registered region"},
    {ErrorCode::IO_ERROR, "IO_ERROR", "Disk or io_uring I/O failed"},
    {ErrorCode::NETWORK_ERROR, "NETWORK_ERROR", "Generic network failure"},
    {ErrorCode::CONNECTION_FAILED, "CONNECTION_FAILED", "Could not connect to the remote endpoint"},
    {ErrorCode::TRANSFER_FAILED, "TRANSFER_FAILED", "Data transfer failed"},
    {ErrorCode::UCX_ERROR, "UCX_ERROR", "Transport library error (legacy code; fabric errors use FABRIC_ERROR)"},
    {ErrorCode::INVALID_ADDRESS, "INVALID_ADDRESS", "Address is malformed"},
    {ErrorCode::REMOTE_ENDPOINT_ERROR, "REMOTE_ENDPOINT_ERROR", "Remote endpoint reported an error"},
    {ErrorCode::RPC_FAILED, "RPC_FAILED", "RPC call failed"},
    {ErrorCode::FABRIC_ERROR, "FABRIC_ERROR", "GPU fabric (peer mapping / NVLink / multicast) error"},
    {ErrorCode::ETCD_ERROR, "ETCD_ERROR", "Coordination store error"},
    {ErrorCode::ETCD_KEY_NOT_FOUND, "ETCD_KEY_NOT_FOUND", "Coordination key not found"},
    {ErrorCode::ETCD_TRANSACTION_FAILED, "ETCD_TRANSACTION_FAILED", "Coordination transaction compare failed"},
    {ErrorCode::ETCD_LEASE_ERROR, "ETCD_LEASE_ERROR", "Lease is unknown or expired"},
    {ErrorCode::ETCD_WATCH_ERROR, "ETCD_WATCH_ERROR", "Watch could not be established"},
    {ErrorCode::LEADER_ELECTION_FAILED, "LEADER_ELECTION_FAILED", "Leader election campaign failed"},
    {ErrorCode::SERVICE_REGISTRATION_FAILED, "SERVICE_REGISTRATION_FAILED", "Service registration failed"},
    {ErrorCode::NOT_LEADER, "NOT_LEADER", "This keystone is not the elected leader"},
    {ErrorCode::OBJECT_NOT_FOUND, "OBJECT_NOT_FOUND", "Object key is unknown or expired"},
    {ErrorCode::OBJECT_ALREADY_EXISTS, "OBJECT_ALREADY_EXISTS", "Object key already exists"},
    {ErrorCode::INVALID_KEY, "INVALID_KEY", "Object key is empty or malformed"},
    {ErrorCode::INVALID_WORKER, "INVALID_WORKER", "Worker id is unknown"},
    {ErrorCode::WORKER_NOT_READY, "WORKER_NOT_READY", "Worker has not finished initialising"},
    {ErrorCode::NO_COMPLETE_WORKER, "NO_COMPLETE_WORKER", "No replica of the object is complete and alive"},
    {ErrorCode::DATA_CORRUPTION, "DATA_CORRUPTION", "Stored data is corrupted"},
    {ErrorCode::CHECKSUM_MISMATCH, "CHECKSUM_MISMATCH", "Digest of transferred data does not match"},
    {ErrorCode::OBJECT_NOT_READY, "OBJECT_NOT_READY", "Object put is still pending"},
    {ErrorCode::CLIENT_ERROR, "CLIENT_ERROR", "Generic client failure"},
    {ErrorCode::CLIENT_NOT_FOUND, "CLIENT_NOT_FOUND", "Client session is unknown"},
    {ErrorCode::CLIENT_ALREADY_EXISTS, "CLIENT_ALREADY_EXISTS", "Client session already registered"},
    {ErrorCode::CLIENT_DISCONNECTED, "CLIENT_DISCONNECTED", "Client is disconnected"},
    {ErrorCode::SESSION_EXPIRED, "SESSION_EXPIRED", "Client session TTL expired"},
    {ErrorCode::INVALID_CLIENT_STATE, "INVALID_CLIENT_STATE", "Client is in the wrong state"},
    {ErrorCode::ACCESS_DENIED, "ACCESS_DENIED", "Cluster token missing or wrong, or key outside the tenant's grants"},
    {ErrorCode::QUOTA_EXCEEDED, "QUOTA_EXCEEDED", "Tenant budget (bytes or objects) exceeded"},
    {ErrorCode::CONFIG_ERROR, "CONFIG_ERROR", "Generic configuration failure"},
    {ErrorCode::INVALID_CONFIGURATION, "INVALID_CONFIGURATION", "Configuration is inconsistent"},
    {ErrorCode::INVALID_PARAMETERS, "INVALID_PARAMETERS", "Parameters are invalid"},
    {ErrorCode::MISSING_REQUIRED_FIELD, "MISSING_REQUIRED_FIELD", "A required configuration field is missing"},
    {ErrorCode::VALUE_OUT_OF_RANGE, "VALUE_OUT_OF_RANGE", "A configuration value is out of range"},
};

const Entry* find(ErrorCode c) noexcept {
  for (const auto& e : kTable)
    if (e.code == c) return &e;
  return nullptr;
}
}  // namespace

std::string_view to_string(ErrorCode c) noexcept {
  const Entry* e = find(c);
  return e ? e->name : std::string_view("UNKNOWN_ERROR");
}

std::string_view get_error_description(ErrorCode c) noexcept {
  const Entry* e = find(c);
  return e ? e->desc : std::string_view("Unknown error code");
}

std::string_view to_string(Domain d) noexcept {
  switch (d) {
    case Domain::SUCCESS: return "SUCCESS";
    case Domain::SYSTEM: return "SYSTEM";
    case Domain::STORAGE: return "STORAGE";
    case Domain::NETWORK: return "NETWORK";
    case Domain::COORDINATION: return "COORDINATION";
    case Domain::DATA: return "DATA";
    case Domain::CLIENT: return "CLIENT";
    case Domain::CONFIG: return "CONFIG";
  }
  return "UNKNOWN";
}

}  // namespace bb
