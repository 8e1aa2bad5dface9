#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3r_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3r_tests.log
tail -6 $O/r3r_tests.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r3r_latency.json > $O/r3r_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)_B1_auto|sgld it|chains" $O/r3r_latency.log
