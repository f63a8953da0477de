#!/bin/bash
N=${1:-8}
OUT=gpurun_out/r2/n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29602 bench.py --gpus $N --config repl3 > $OUT/repl3.json 2> $OUT/repl3.err
tail -c 1800 $OUT/repl3.json; echo; grep -v "^\*\*\*\|OMP_NUM\|^$" $OUT/repl3.err | tail -5
