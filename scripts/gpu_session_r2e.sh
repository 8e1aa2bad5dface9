#!/bin/bash
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2e_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2e_tests.log
tail -4 $O/r2e_tests.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r2e_bench_n1.json 2> $O/r2e_bench_n1.err; echo "bench rc=$?"
tail -3 $O/r2e_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
print('f64', {k:(v['value'], v['roofline']['frac'], v['e2e']['value']) for k,v in d['f64'].items()})
print(json.dumps(d['extra'], indent=1)[:3000])
print(d.get('cpu_baseline'))
PY
