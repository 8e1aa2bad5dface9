#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize_properties.py tests/test_gpu_edge_cases.py tests/test_gpu_device_loop.py -m gpu -q -p no:cacheprovider -x > $O/r3q_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3q_tests.log
tail -6 $O/r3q_tests.log
timeout 300 python scripts/probe_clsync.py > $O/r3q_clsync.log 2>&1; cat $O/r3q_clsync.log | tail -21
SGM_NO_CLSYNC=1 timeout 300 python scripts/probe_clsync.py > $O/r3q_noclsync.log 2>&1; cat $O/r3q_noclsync.log | tail -21
