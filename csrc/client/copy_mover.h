// Data mover used by the keystone for tier demotion and re-replication (the
// reference "evicts" by forgetting objects; nothing ever moves between tiers).
//
// make_data_server_mover(): copies one replica to a new placement through the workers' data
// servers.  When source and destination shard live in the same worker process the bytes never
// leave it (D_COPY: backend -> pinned staging -> backend, e.g. GPU slab -> DRAM -> NVMe via
// cudaMemcpyAsync / io_uring on the worker's side streams); otherwise they are relayed.
// Every destination shard gets a fresh digest, and the source digest is verified on the way.
#pragma once
#include "keystone/keystone_service.h"

namespace bb::client {

keystone::CopyMover make_data_server_mover(size_t io_parallelism = 4, int rpc_timeout_ms = 30000);

// Scrub: asks the worker that holds each shard of `copy` to hash it where it lies (D_CHECKSUM: the bytes never cross the
// network) and compares with the digest recorded at put_complete.  OK = every shard matches, CHECKSUM_MISMATCH = bit rot,
// anything else = the worker could not be asked.
keystone::CopyVerifier make_data_server_verifier(int rpc_timeout_ms = 30000);

// The Keystone's side of the reservation protocol, spoken to the workers' data servers (D_RESERVE / D_COMMIT / D_ABORT /
// D_FREE, one request per pool): put_start reserves every shard of the placement at its worker, put_complete commits,
// put_cancel / expiry aborts, removing a COMPLETE object frees.  Reference: the contract of
// include/blackbird/worker/storage/storage_backend.h:46-126, which no service there ever calls.
keystone::ReservationHooks make_data_server_reservation_hooks(int rpc_timeout_ms = 5000);

}  // namespace bb::client
