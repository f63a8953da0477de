#!/bin/bash
# Host data path benchmark (no GPU): bb-coord + bb-keystone + one bb-worker with a 4 GiB DRAM pool on loopback, then
# `bb-bench client` (put + verified get_into a reused buffer, like the reference's clients/benchmark_client.cpp) at
# several object sizes.
#   bench/host_path_bench.sh [--shm] [extra bb-bench flags, e.g. --checksum crc32c|none]
#   --shm : the pool is memfd-backed (shared_memory: true) -> same-host clients move bytes with one-sided memcpy
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
BIN="${BB_BIN_DIR:-$ROOT/bin}"
SHM=false
if [ "${1:-}" = "--shm" ]; then SHM=true; shift; fi
D="$(mktemp -d)"; CP=$((20000 + RANDOM % 20000)); RP=$((CP + 1)); HP=$((CP + 2))
cleanup() { for f in w ks coord; do [ -f "$D/$f.pid" ] && kill "$(cat "$D/$f.pid")" 2>/dev/null || true; done; rm -rf "$D"; }
trap cleanup EXIT
"$BIN/bb-coord" --listen 127.0.0.1:$CP > "$D/coord.log" 2>&1 & echo $! > "$D/coord.pid"
sleep 0.3
"$BIN/bb-keystone" "$ROOT/configs/keystone.yaml" --coord-endpoints 127.0.0.1:$CP --listen-address 127.0.0.1:$RP --http-port $HP > "$D/ks.log" 2>&1 & echo $! > "$D/ks.pid"
sleep 0.3
cat > "$D/w.yaml" <<Y
worker: {worker_id: "w0", node_id: "n0", interconnects: ["tcp"], lease_ttl_sec: 10, heartbeat_interval_sec: 3}
storage_pools:
  - {pool_id: "ram0", storage_class: "RAM_CPU", size_bytes: 4_GB, shared_memory: $SHM}
Y
"$BIN/bb-worker" --config "$D/w.yaml" --coord-endpoints 127.0.0.1:$CP > "$D/w.log" 2>&1 & echo $! > "$D/w.pid"
sleep 1
for sz in 1024 65536 1048576 16777216 134217728; do
  it=20; [ $sz -le 65536 ] && it=2000; [ $sz -ge 134217728 ] && it=6
  "$BIN/bb-bench" client --keystone 127.0.0.1:$RP --size $sz --iterations $it "$@" | tail -1
done
