#!/bin/bash
O=gpurun_out
P=stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
timeout 600 python -m pytest tests/test_gpu_kernel_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_statistical_models.py tests/test_gpu_device_loop.py tests/test_gpu_fullsize_properties.py -m gpu -q -p no:cacheprovider > $O/r2r_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2r_tests.log
tail -3 $O/r2r_tests.log
timeout 300 python scripts/probe_small_shape.py > $O/r2r_small_shape.log 2>&1; echo "rc=$?"
cat $O/r2r_small_shape.log
SGM_LIB_PATH=$PWD/$P/libsgmpf_s256x4.so timeout 300 python scripts/probe_small_shape.py > $O/r2r_small_shape_256x4.log 2>&1; echo "rc=$?"
cat $O/r2r_small_shape_256x4.log
