#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2u_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2u_tests.log
tail -4 $O/r2u_tests.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r2u_latency.json > $O/r2u_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)|sgld it|chains" $O/r2u_latency.log
