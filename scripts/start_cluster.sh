#!/bin/bash
# Starts a local blackbird_b200 cluster: bb-coord -> bb-keystone -> N bb-workers -> smoke test.
# (Role of the reference's scripts/start_cluster.sh: etcd -> keystone -> worker -> smoke.)
#   scripts/start_cluster.sh [-n WORKERS] [-d RUN_DIR] [--gpu] [--ha] [--secure]     env: BB_COORD_PORT BB_RPC_PORT BB_HTTP_PORT
#   --secure: a fresh cluster token (RUN_DIR/token, mode 600) gates every RPC server and keys AES-256-GCM on every frame;
#             clients need  BB_AUTH_TOKEN="$(cat RUN_DIR/token)" BB_ENCRYPT_TRANSPORT=1
#   --ha: durable bb-coord (log + snapshots under RUN_DIR/coord-data) and a Keystone pair (second one on RPC_PORT+10 /
#         HTTP_PORT+10); the elected leader serves, the standby takes over from the metadata log; clients get both endpoints.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
BIN="$ROOT/bin"
N=1; RUN_DIR="${TMPDIR:-/tmp}/blackbird_b200_cluster"; GPU=0; HA=0; SECURE=0
while [ $# -gt 0 ]; do case "$1" in -n) N="$2"; shift 2;; -d) RUN_DIR="$2"; shift 2;; --gpu) GPU=1; shift;; --ha) HA=1; shift;; --secure) SECURE=1; shift;; *) echo "unknown arg $1"; exit 2;; esac; done
COORD_PORT="${BB_COORD_PORT:-2379}"; RPC_PORT="${BB_RPC_PORT:-9090}"; HTTP_PORT="${BB_HTTP_PORT:-9091}"
mkdir -p "$RUN_DIR"
if [ "$SECURE" = 1 ]; then
  ( umask 077; head -c 24 /dev/urandom | base64 | tr -d '\n' > "$RUN_DIR/token" )
  export BB_AUTH_TOKEN="$(cat "$RUN_DIR/token")" BB_ENCRYPT_TRANSPORT=1
fi
for b in bb-coord bb-keystone bb-worker bb-cli; do [ -x "$BIN/$b" ] || { echo "missing $BIN/$b (run: python build.py)"; exit 1; }; done
port_free() { ! (exec 3<>"/dev/tcp/127.0.0.1/$1") 2>/dev/null; }
for p in "$COORD_PORT" "$RPC_PORT" "$HTTP_PORT"; do port_free "$p" || { echo "port $p is in use"; exit 1; }; done
wait_port() { for _ in $(seq 1 100); do if ! port_free "$1"; then return 0; fi; sleep 0.1; done; echo "timeout waiting for port $1"; return 1; }

COORD_ARGS=(); KS_ARGS=(); KEYSTONES="127.0.0.1:$RPC_PORT"
if [ "$HA" = 1 ]; then
  RPC2=$((RPC_PORT + 10)); HTTP2=$((HTTP_PORT + 10))
  for p in "$RPC2" "$HTTP2"; do port_free "$p" || { echo "port $p is in use"; exit 1; }; done
  COORD_ARGS=(--data-dir "$RUN_DIR/coord-data"); KS_ARGS=(--enable-ha); KEYSTONES="127.0.0.1:$RPC_PORT,127.0.0.1:$RPC2"
fi
"$BIN/bb-coord" --listen "127.0.0.1:$COORD_PORT" "${COORD_ARGS[@]}" > "$RUN_DIR/coord.log" 2>&1 & echo $! > "$RUN_DIR/coord.pid"
wait_port "$COORD_PORT"
"$BIN/bb-keystone" "$ROOT/configs/keystone.yaml" --coord-endpoints "127.0.0.1:$COORD_PORT" --listen-address "127.0.0.1:$RPC_PORT" \
    --http-port "$HTTP_PORT" "${KS_ARGS[@]}" > "$RUN_DIR/keystone.log" 2>&1 & echo $! > "$RUN_DIR/keystone.pid"
wait_port "$RPC_PORT"
if [ "$HA" = 1 ]; then
  "$BIN/bb-keystone" "$ROOT/configs/keystone.yaml" --coord-endpoints "127.0.0.1:$COORD_PORT" --listen-address "127.0.0.1:$RPC2" \
      --http-port "$HTTP2" --enable-ha > "$RUN_DIR/keystone2.log" 2>&1 & echo $! > "$RUN_DIR/keystone2.pid"
  wait_port "$RPC2"
fi
for i in $(seq 0 $((N - 1))); do
  if [ "$GPU" = 1 ]; then
    cat > "$RUN_DIR/worker$i.yaml" <<YAML
worker: {worker_id: "worker-gpu$i", node_id: "gpu$i", interconnects: ["nvlink", "tcp"], fabric_domain: "nvswitch-0", lease_ttl_sec: 10, heartbeat_interval_sec: 3}
storage_pools:
  - {pool_id: "hbm$i", storage_class: "RAM_GPU", size_bytes: 4_GB, gpu_device_id: $i}
  - {pool_id: "dram$i", storage_class: "RAM_CPU", size_bytes: 1_GB}
YAML
  else
    cat > "$RUN_DIR/worker$i.yaml" <<YAML
worker: {worker_id: "worker-$i", node_id: "node-$i", interconnects: ["tcp"], lease_ttl_sec: 10, heartbeat_interval_sec: 3}
storage_pools:
  - {pool_id: "ram$i", storage_class: "RAM_CPU", size_bytes: 256_MB}
  - {pool_id: "nvme$i", storage_class: "NVME", size_bytes: 256_MB, mount_path: "$RUN_DIR/nvme$i"}
YAML
  fi
  "$BIN/bb-worker" --config "$RUN_DIR/worker$i.yaml" --coord-endpoints "127.0.0.1:$COORD_PORT" > "$RUN_DIR/worker$i.log" 2>&1 & echo $! > "$RUN_DIR/worker$i.pid"
done
sleep 0.5
"$BIN/bb-cli" --keystone "$KEYSTONES" smoke --size 1024
"$BIN/bb-cli" metrics --http "127.0.0.1:$HTTP_PORT" | grep -E "^bb_(workers|memory_pools|objects) "
echo "cluster is up: coord=$COORD_PORT keystone=$KEYSTONES metrics=http://127.0.0.1:$HTTP_PORT/metrics run_dir=$RUN_DIR"
[ "$SECURE" = 1 ] && echo "secure cluster: clients need BB_AUTH_TOKEN=\"\$(cat $RUN_DIR/token)\" BB_ENCRYPT_TRANSPORT=1"
echo "stop with: scripts/stop_cluster.sh -d $RUN_DIR"
