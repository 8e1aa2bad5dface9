#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2q_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2q_tests.log
tail -6 $O/r2q_tests.log
timeout 300 python scripts/probe_small_shape.py > $O/r2q_small_shape.log 2>&1; echo "rc=$?"
cat $O/r2q_small_shape.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $O/r2q_bench.log 2>&1; tail -1 $O/r2q_bench.log | cut -c1-600
