#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3s_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3s_tests.log
tail -3 $O/r3s_tests.log
timeout 300 python scripts/probe_clsync.py > $O/r3s_clsync.log 2>&1; cat $O/r3s_clsync.log | tail -21
timeout 900 python scripts/probe_latency.py --quick --json $O/r3s_latency.json > $O/r3s_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)_B1_auto|sgld it.*(8192|65536)|chains" $O/r3s_latency.log
