#!/bin/bash
# Last 8-GPU confirmation of the final tree: headline line + quick size sweep.
N=${1:-8}
OUT=gpurun_out/r2/n${N}_last
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
timeout 250 $TR --master-port 29632 bench.py --gpus $N --config sweep --quick > $OUT/sweep.json 2> $OUT/sweep.err
python - <<PY
import json
def load(p):
    t=open(p).read(); return json.loads(t[t.find('{"'):t.rfind('}')+1])
b=load('$OUT/bench.json')
print(b['value'], b['ms_per_step'], b['kernel_only'], b['e2e']['value'], b['latency_us_single_object'])
for r in load('$OUT/sweep.json')['sweep']: print(r['size'], r['batch'], r['put_GBps_client'], r['get_GBps_client'], r['put_p50_us'], r['get_p50_us'], r['put_p99_us'], r['get_p99_us'])
PY
