#!/bin/bash
OUT=gpurun_out/r2/n2_sweep
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --config sweep --quick > $OUT/sweep.json 2> $OUT/sweep.err
python - <<'PY'
import json
t=open('gpurun_out/r2/n2_sweep/sweep.json').read(); d=json.loads(t[t.find('{"'):t.rfind('}')+1])
for r in d['sweep']: print(r['size'], r['batch'], 'kernel', r['put_GBps_kernel'], r['get_GBps_kernel'], 'client', r['put_GBps_client'], r['get_GBps_client'], 'p50', r['put_p50_us'], r['get_p50_us'])
print(d['clocks'])
PY
