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

}  // namespace bb::client
