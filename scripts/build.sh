#!/bin/bash
# Rebuild libsgmpf.so (sm_100a) from any cwd through the package's own recipe (three translation units compiled in
# parallel, then linked); prints ptxas resource lines for the kernels matching $1.  Extra nvcc flags: SGM_NVCC_EXTRA.
cd "$(dirname "$0")/.." || exit 1
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
PYTHONPATH="$PKG" python -m sgmcmc_ssm_b200.build --force -v > /tmp/nvcc_build.out 2> /tmp/nvcc_build.log
rc=$?
grep -i "error" /tmp/nvcc_build.log | head -20
if [ -n "$1" ]; then grep -A2 "$1" /tmp/nvcc_build.log | grep -v "^--" | head -12; fi
exit $rc
