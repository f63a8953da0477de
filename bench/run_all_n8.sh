#!/bin/bash
# Everything that needs the 8-GPU box in one gpurun call (8x GPU-minutes: keep it short).  Outputs -> gpurun_out/r2/n8/
N=${1:-8}
OUT=gpurun_out/r2/n$N
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 120 bin/bb-p2p-probe --gpus $N --quick --mib 512 > $OUT/p2p_probe.jsonl 2> $OUT/p2p_probe.err
timeout 300 $TR --master-port 29601 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
timeout 200 $TR --master-port 29602 bench.py --gpus $N --config repl3 > $OUT/repl3.json 2> $OUT/repl3.err
timeout 150 $TR --master-port 29603 bench.py --gpus $N --config fanout > $OUT/fanout.json 2> $OUT/fanout.err
timeout 250 $TR --master-port 29604 bench.py --gpus $N --config sweep --quick > $OUT/sweep.json 2> $OUT/sweep.err
timeout 150 $TR --master-port 29605 bench/nvls_test.py --mode broadcast > $OUT/nvls_broadcast.txt 2>&1
for f in bench repl3 fanout sweep; do echo "== $f"; tail -c 1500 $OUT/$f.json; echo; tail -3 $OUT/$f.err; done
tail -5 $OUT/nvls_broadcast.txt
grep -E "ring|mc_" $OUT/p2p_probe.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('%-10s %-12s %7.1f %7.1f  %s' % (d['method'], d['pattern'], d['GBps_per_gpu_min'], d.get('delivered_GBps_per_writer', 0), d['cfg']))
" | head -80
