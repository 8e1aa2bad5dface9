#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3k_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3k_tests.log
tail -4 $O/r3k_tests.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r3k_latency.json > $O/r3k_latency.log 2>&1
grep -E "svm_N(1000|2048|4096)_B(1|8)_(auto|small)|sgld it|chains" $O/r3k_latency.log
