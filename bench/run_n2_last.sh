#!/bin/bash
# Last checks of the final tree: full GPU test-suite on GPU 0, then the 2-GPU headline line.
OUT=gpurun_out/r2/n2_last
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/pytest_gpu.txt
python - <<'PY'
import json
t=open('gpurun_out/r2/n2_last/bench.json').read(); d=json.loads(t[t.find('{"'):t.rfind('}')+1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['mode'], d['latency_us_single_object'])
PY
