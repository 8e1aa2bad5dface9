#!/bin/bash
# what the driver runs at round end, on the final library: GPU suite, smoke, reference arm, bench
O=gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > $O/r3m_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3m_tests.log
tail -3 $O/r3m_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r3m_bench_ref.log 2> $O/r3m_bench_ref.err; echo "ref rc=$?"; tail -1 $O/r3m_bench_ref.log | cut -c1-200
timeout 1500 python bench.py > $O/r3m_bench_n1.log 2> $O/r3m_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3m_bench_n1.log') if l.startswith('{')][-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
print('f64 native',d['f64']['native']['value'],d['f64']['native']['roofline']['frac'],'v32',d['f64']['variates_f32']['value'])
print('cpu_baseline',d['cpu_baseline']['value'],d['cpu_baseline']['kind'])
for k,v in d['extra'].items():
    if isinstance(v,(int,float,bool)): print(k, v)
PY
