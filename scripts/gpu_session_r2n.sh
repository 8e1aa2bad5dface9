#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize_properties.py tests/test_gpu_edge_cases.py -m gpu -q -p no:cacheprovider > $O/r2n_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2n_tests.log
tail -5 $O/r2n_tests.log
timeout 300 python scripts/sanitizer_target.py > $O/r2n_sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/r2n_sanitizer_plain.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitizer_target.py > $O/r2n_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -8 $O/r2n_memcheck.log | cut -c1-200
