#!/bin/bash
# A/B: TMA (cp.async.bulk) staging of the raw parent-CDF window in the FAST step kernel
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
SGM_LIB_PATH=$P/libsgmpf_tma.so timeout 900 python -m pytest tests/test_gpu_statistical_models.py tests/test_gpu_fullsize_properties.py tests/test_gpu_fullsize_window.py -m gpu -q -p no:cacheprovider > $O/r2s_tests_tma.log 2>&1; echo "pytest rc=$?" >> $O/r2s_tests_tma.log
tail -8 $O/r2s_tests_tma.log
for rep in 1 2; do
timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $O/r2s_bench_base_$rep.log 2>&1; tail -1 $O/r2s_bench_base_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('base', d['value'], d['roofline']['frac'])"
SGM_LIB_PATH=$P/libsgmpf_tma.so timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $O/r2s_bench_tma_$rep.log 2>&1; tail -1 $O/r2s_bench_tma_$rep.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tma ', d['value'], d['roofline']['frac'])"
done
