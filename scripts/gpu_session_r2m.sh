#!/bin/bash
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2m_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2m_tests.log
tail -4 $O/r2m_tests.log
python scripts/probe_cluster.py 2>&1 | cut -c1-60 | tee $O/r2m_cluster.log
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/r2m_bench_n1.json 2> $O/r2m_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
print('f64', {k:(v['value'], v['roofline']['frac'], v['e2e']['value']) for k,v in d['f64'].items()})
print({k:v for k,v in d['extra'].items() if not isinstance(v, dict)})
print(d['extra']['config5_strong']); print(d['extra']['config4_chains']['N1000'], d['extra']['config4_chains']['N10000'])
PY
