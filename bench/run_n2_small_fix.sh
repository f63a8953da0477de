#!/bin/bash
OUT=gpurun_out/r2/small_fix
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_stack.py -x -q -k "small or mailbox or fused_digest or batch_of_small" > $OUT/pytest_small.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_small.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench/lat256.py > $OUT/lat256.jsonl 2> $OUT/lat256.err
tail -3 $OUT/pytest_small.txt
python - <<'PY'
import json
for l in open('gpurun_out/r2/small_fix/lat256.jsonl'):
    l=l.strip()
    if not l.startswith('{'): continue
    d=json.loads(l)
    for sect in ('remote','local'):
        for r in d[sect]: print(d['rank'], sect, r['size'], 'put', round(r['put_p50_us'],1), 'get', round(r['get_p50_us'],1), r['engine_paths'])
PY
