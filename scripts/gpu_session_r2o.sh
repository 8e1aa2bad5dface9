#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_statistical_models.py tests/test_gpu_device_loop.py tests/test_gpu_fullsize_properties.py tests/test_gpu_resampling_schemes.py -m gpu -q -p no:cacheprovider > $O/r2o_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2o_tests.log
tail -6 $O/r2o_tests.log
timeout 300 python scripts/probe_small_shape.py > $O/r2o_small_shape.log 2>&1; echo "rc=$?"
cat $O/r2o_small_shape.log
