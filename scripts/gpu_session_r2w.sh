#!/bin/bash
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
timeout 900 python -m pytest tests/test_gpu_fullsize_properties.py tests/test_gpu_kernel_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_bign_parity.py tests/test_gpu_fullsize_window.py -m gpu -q -p no:cacheprovider > $O/r2w_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2w_tests.log
tail -4 $O/r2w_tests.log
SGM_LIB_PATH=$P/libsgmpf_cooptime.so timeout 300 python scripts/probe_coop_timing.py 2>&1 | grep -v "^$" | tail -8
timeout 900 python scripts/probe_latency.py --quick --json $O/r2w_latency.json > $O/r2w_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)|sgld it.*(8192|65536)" $O/r2w_latency.log
