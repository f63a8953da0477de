#!/bin/bash
N=${1:-4}
OUT=gpurun_out/r2/n${N}_final
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
head -c 500 $OUT/bench.json; echo; grep -o '"latency_us_single_object": {.*}}, "roofline' $OUT/bench.json | head -c 800
