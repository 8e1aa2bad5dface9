#!/bin/bash
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
for rep in 1 2; do
for v in base f64c5 f64c6; do
  if [ $v = base ]; then unset SGM_LIB_PATH; else export SGM_LIB_PATH=$P/libsgmpf_$v.so; fi
  echo "== $v rep $rep"
  timeout 600 python scripts/probe_models.py --dtype f64 --variates native --pf poyiadjis_N,nemeth 2>&1 | grep -E "poyiadjis_N|nemeth" | cut -c1-110
done
done
unset SGM_LIB_PATH
for v in base f64c6; do
  if [ $v = base ]; then unset SGM_LIB_PATH; else export SGM_LIB_PATH=$P/libsgmpf_$v.so; fi
  echo "== $v f32 variates"
  timeout 600 python scripts/probe_models.py --dtype f64 --variates f32 --pf poyiadjis_N 2>&1 | grep -E "poyiadjis_N" | cut -c1-110
done
