#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3b_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3b_tests.log
tail -4 $O/r3b_tests.log
timeout 900 python scripts/probe_latency.py --quick --json $O/r3b_latency.json > $O/r3b_latency.log 2>&1
grep -E "svm_N(4096|8192|16384|65536)|sgld it|chains" $O/r3b_latency.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > $O/r3b_bench.log 2>&1; tail -1 $O/r3b_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['frac'])"
