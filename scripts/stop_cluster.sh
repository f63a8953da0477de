#!/bin/bash
# Stops the processes started by start_cluster.sh using the recorded PIDs (never by pattern).
set -uo pipefail
RUN_DIR="${TMPDIR:-/tmp}/blackbird_b200_cluster"
while [ $# -gt 0 ]; do case "$1" in -d) RUN_DIR="$2"; shift 2;; *) shift;; esac; done
for f in "$RUN_DIR"/worker*.pid "$RUN_DIR/keystone2.pid" "$RUN_DIR/keystone.pid" "$RUN_DIR/coord.pid"; do
  [ -f "$f" ] || continue
  pid="$(cat "$f")"
  if kill -0 "$pid" 2>/dev/null; then kill "$pid"; for _ in $(seq 1 50); do kill -0 "$pid" 2>/dev/null || break; sleep 0.1; done; kill -0 "$pid" 2>/dev/null && kill -9 "$pid"; fi
  rm -f "$f"
  echo "stopped $(basename "$f" .pid) ($pid)"
done
