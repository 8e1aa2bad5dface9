#!/bin/bash
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2g_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2g_tests.log
tail -5 $O/r2g_tests.log
timeout 900 python scripts/probe_latency.py --json $O/r2g_latency.json > $O/r2g_latency.log 2>&1; echo "rc=$?" >> $O/r2g_latency.log
cat $O/r2g_latency.log | tail -80
