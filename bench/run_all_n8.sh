#!/bin/bash
# Everything that needs the 8-GPU box in one gpurun call (8x GPU-minutes: keep it short).  Outputs -> gpurun_out/r2/n8/
N=${1:-8}
OUT=gpurun_out/r2/n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 240 python -m pytest tests/test_multi_gpu.py -x -q > $OUT/pytest_multi_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_multi_gpu.txt
timeout 300 $TR --master-port 29601 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
timeout 200 $TR --master-port 29602 bench.py --gpus $N --config repl3 > $OUT/repl3.json 2> $OUT/repl3.err
timeout 150 $TR --master-port 29603 bench.py --gpus $N --config fanout > $OUT/fanout.json 2> $OUT/fanout.err
timeout 250 $TR --master-port 29604 bench.py --gpus $N --config sweep --quick > $OUT/sweep.json 2> $OUT/sweep.err
timeout 150 $TR --master-port 29605 bench/nvls_test.py --mode broadcast > $OUT/nvls_broadcast.txt 2>&1
for f in bench repl3 fanout sweep; do echo "== $f"; tail -c 1500 $OUT/$f.json; echo; tail -3 $OUT/$f.err; done
tail -5 $OUT/nvls_broadcast.txt
tail -3 $OUT/pytest_multi_gpu.txt
