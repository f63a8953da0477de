// Fault injection (the reference has no hooks at all).
//
// Named faults are armed from the environment (BB_FAULT="drop_heartbeat,corrupt_write:2,delay_rpc_ms=20")
// or programmatically (tests).  `name:count` fires `count` times and disarms itself; `name=value` carries a
// number (e.g. a delay in ms); a bare name stays on until cleared.  Injection points:
//   drop_heartbeat      worker stops refreshing its lease            (worker_service.cpp)
//   corrupt_write       data server flips a byte after a D_WRITE     (worker_service.cpp)  -> CHECKSUM_MISMATCH on read
//   fail_data_read      data server answers D_READ with IO_ERROR     (worker_service.cpp)  -> replica fail-over
//   fail_data_write     data server answers D_WRITE with IO_ERROR    (worker_service.cpp)  -> put_cancel
//   fail_put_complete   keystone rejects put_complete                (keystone_service.cpp)
//   delay_rpc_ms=<n>    every keystone RPC handler sleeps n ms       (rpc_service.cpp)     -> client time-outs
#pragma once
#include <cstdint>
#include <string>

namespace bb::fault {

// True when `name` is armed; consumes one shot of a counted fault.
bool fire(const char* name);
// Value attached to an armed fault (does not consume it); `def` when not armed.
int64_t value(const char* name, int64_t def = 0);
// count < 0: until cleared.
void arm(const std::string& name, int64_t value = 1, int64_t count = -1);
void disarm(const std::string& name);
void clear();
// Parses a BB_FAULT-style spec; returns the number of faults armed.
size_t arm_from_spec(const std::string& spec);
bool any_armed();

}  // namespace bb::fault
