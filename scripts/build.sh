#!/bin/bash
# Rebuild libsgmpf.so (sm_100a) from any cwd; prints ptxas resource lines for the kernels matching $1.
cd "$(dirname "$0")/.." || exit 1
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v \
  -o "$PKG/libsgmpf.so" "$PKG/csrc/sgmpf.cu" > /tmp/nvcc_build.log 2>&1
rc=$?
grep -i "error" /tmp/nvcc_build.log | head -20
if [ -n "$1" ]; then grep -A2 "$1" /tmp/nvcc_build.log | grep -v "^--" | head -12; fi
exit $rc
