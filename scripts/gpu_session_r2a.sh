#!/bin/bash
# round-2 GPU session A: full -m gpu suite with the new reference-pinned tests, fast-mode probes, f64 shape A/B
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > $O/r2a_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2a_tests.log
tail -5 $O/r2a_tests.log
timeout 300 python scripts/probe_models.py --json $O/r2a_probe_f32.json > $O/r2a_probe_f32.log 2>&1
timeout 300 python scripts/probe_models.py --dtype f64 --json $O/r2a_probe_f64.json > $O/r2a_probe_f64.log 2>&1
timeout 300 python scripts/probe_models.py --dtype f64 --variates f32 --json $O/r2a_probe_f64v32.json > $O/r2a_probe_f64v32.log 2>&1
for v in f64c4 f64c6; do
  SGM_LIB_PATH=$PWD/$PKG/libsgmpf_$v.so timeout 200 python scripts/probe_models.py --dtype f64 --models svm,garch --pf poyiadjis_N > $O/r2a_probe_$v.log 2>&1
  SGM_LIB_PATH=$PWD/$PKG/libsgmpf_$v.so timeout 200 python scripts/probe_models.py --dtype f64 --variates f32 --models svm,garch --pf poyiadjis_N >> $O/r2a_probe_$v.log 2>&1
done
cat $O/r2a_probe_*.log
