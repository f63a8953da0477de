#!/bin/bash
# Re-validation of the tree after the last changes (1 GPU): GPU tests, headline line, single-object latency, small-batch sweep.
OUT=gpurun_out/r2/n1_recheck
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.txt 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench/latency_probe.py > $OUT/latency.json 2> $OUT/latency.err
timeout 300 python bench.py --config sweep --quick > $OUT/sweep.json 2> $OUT/sweep.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.txt 2>&1
timeout 100 python bench.py --impl reference > $OUT/reference_arm.json 2>&1
tail -3 $OUT/pytest_gpu.txt; head -c 400 $OUT/bench.json; echo; grep -o '"latency_us_single_object": {[^}]*}[^}]*}[^}]*}[^}]*}' $OUT/bench.json; tail -2 $OUT/smoke.txt; cat $OUT/reference_arm.json | head -c 400; echo
python - <<'PY'
import json
t=open('gpurun_out/r2/n1_recheck/latency.json').read()
d=json.loads(t[t.find('{'):t.rfind('}')+1])
for r in d['engine']:
    if r['algo']=='XXH3' and r['mailbox']: print('engine mailbox', r['size'], r['call_p50_us'], r['call_p99_us'])
for r in d['client']['small_path_on']: print('client', r['size'], round(r['put_p50_us'],1), round(r['get_p50_us'],1))
t=open('gpurun_out/r2/n1_recheck/sweep.json').read()
d=json.loads(t[t.find('{"'):t.rfind('}')+1])
for r in d['sweep']: print(r['size'], r['batch'], r['put_GBps_client'], r['get_GBps_client'], r['put_p50_us'], r['get_p50_us'], r.get('rank0_client_phases'))
PY
