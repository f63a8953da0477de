#!/bin/bash
# Final-tree checks that need 2 GPUs: smoke() (multi-GPU branch), the torchrun GPU tests, the headline line.
N=${1:-2}
OUT=gpurun_out/r2/n${N}_final
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.txt 2>&1; echo "smoke exit $?" >> $OUT/smoke.txt
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q > $OUT/pytest_multi_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi_gpu.txt
timeout 300 $TR --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -4 $OUT/smoke.txt; tail -3 $OUT/pytest_multi_gpu.txt; head -c 600 $OUT/bench.json; echo; grep -o '"clocks": {[^}]*}' $OUT/bench.json
