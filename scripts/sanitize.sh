#!/bin/bash
# Sanitizer runs.  CPU side: BB_SANITIZE=asan|tsan python build.py && pytest (or the tsan'd bb-bench
# control).  GPU side (needs a B200): compute-sanitizer over bench/sanitize_target.py, one tool per run.
#   scripts/sanitize.sh [memcheck|synccheck|racecheck|initcheck] [out_dir]
set -u
TOOL=${1:-memcheck}
OUT=${2:-gpurun_out}
mkdir -p "$OUT"
exec compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 9 python bench/sanitize_target.py > "$OUT/sanitize_$TOOL.log" 2>&1
