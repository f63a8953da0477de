#!/bin/bash
# usage: bench/run_multi.sh N [extra bench.py args]  -- launches bench.py on N GPUs of this node
N=$1; shift
PORT=${MASTER_PORT:-29517}
if [ "$N" = "1" ]; then exec python bench.py --gpus 1 "$@"; fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$PORT" bench.py --gpus "$N" "$@"
