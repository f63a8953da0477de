#!/bin/bash
OUT=gpurun_out/r2/control_last
mkdir -p $OUT
for t in 1 8; do bin/bb-bench control --threads $t --iterations 20; bin/bb-bench control --threads $t --iterations 20 --rpc; done > $OUT/control.jsonl 2>&1
bin/bb-bench control --threads 1 --batch 1 --iterations 50000 >> $OUT/control.jsonl 2>&1
bin/bb-bench devclient --iterations 20 >> $OUT/control.jsonl 2>&1
cat $OUT/control.jsonl
timeout 120 python -m pytest tests/test_gpu_stack.py -x -q -k "full_stack or spill or large_batch" > $OUT/pytest_stack.txt 2>&1; tail -2 $OUT/pytest_stack.txt
