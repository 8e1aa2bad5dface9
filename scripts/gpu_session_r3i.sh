#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3i_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3i_tests.log
tail -4 $O/r3i_tests.log
timeout 300 python scripts/probe_small_shape.py > $O/r3i_small_shape.log 2>&1; grep -E "B=1 |B=148|B=296|sgld" $O/r3i_small_shape.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r3i_latency.json > $O/r3i_latency.log 2>&1
grep -E "svm_N(256|1000|1024|2048)_B(1|64|296)_auto|sgld it|chains" $O/r3i_latency.log
timeout 1500 python bench.py > $O/r3i_bench_n1.log 2> $O/r3i_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3i_bench_n1.log') if l.startswith('{')][-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
print('f64 native',d['f64']['native']['value'],d['f64']['native']['roofline']['frac'])
for k,v in d['extra'].items():
    if 'sgld' in k or 'minibatch1' in k: print(k, v)
print(d['extra']['config4_chains']['N1000'])
PY
