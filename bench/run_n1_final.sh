#!/bin/bash
# 1-GPU validation of the final tree: GPU tests, headline bench, size sweep, tier spill, control plane, sanitizers.
OUT=gpurun_out/r2/n1_final
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 400 python bench.py --config sweep > $OUT/sweep.json 2> $OUT/sweep.err
timeout 300 python bench.py --config spill > $OUT/spill.json 2> $OUT/spill.err
for t in 1 8; do bin/bb-bench control --threads $t --iterations 20; bin/bb-bench control --threads $t --iterations 20 --rpc; done > $OUT/control.jsonl 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python bench/sanitize_target.py > $OUT/memcheck.log 2>&1
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python bench/sanitize_target.py > $OUT/synccheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 400 python bench/sanitize_target.py > $OUT/racecheck.log 2>&1
tail -3 $OUT/pytest_gpu.txt; tail -c 600 $OUT/bench.json; echo; tail -c 1500 $OUT/spill.json; echo; cat $OUT/control.jsonl
for f in memcheck synccheck racecheck; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target" $OUT/$f.log | tail -3; done
