#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3a_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3a_tests.log
tail -4 $O/r3a_tests.log
timeout 1500 python bench.py > $O/r3a_bench_n1.log 2> $O/r3a_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3a_bench_n1.log') if l.startswith('{')][-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
print('f64 native',d['f64']['native']['value'],d['f64']['native']['roofline']['frac'],'v32',d['f64']['variates_f32']['value'],d['f64']['variates_f32']['roofline']['frac'])
PY
timeout 600 python scripts/probe_models.py --dtype f64 --variates native --json $O/r3a_models_f64.json 2>&1 | cut -c1-112 | tail -9
