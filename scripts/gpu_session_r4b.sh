#!/bin/bash
# A/B: evict-first (st.global.cs / ld.global.cs) hints on the record stores / gathers of the step kernel
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
for rep in 1 2; do
for v in base sts stsl; do
  if [ $v = base ]; then unset SGM_LIB_PATH; else export SGM_LIB_PATH=$P/libsgmpf_$v.so; fi
  timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $O/r4b_bench_${v}_$rep.log 2>&1; tail -1 $O/r4b_bench_${v}_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['frac'])"
done
done
