#!/bin/bash
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2l_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2l_tests.log
tail -6 $O/r2l_tests.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r2l_latency.json > $O/r2l_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)|sgld it|chains" $O/r2l_latency.log
SGM_NO_COOP=1 timeout 900 python scripts/probe_latency.py --quick > $O/r2l_latency_nocoop.log 2>&1
echo "== SGM_NO_COOP"; grep -E "svm_N(4096|8192|16384|65536)|sgld it|chains" $O/r2l_latency_nocoop.log
