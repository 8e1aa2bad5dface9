#!/bin/bash
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
for rep in 1 2; do
for v in base ws8 ws7; do
  if [ $v = base ]; then unset SGM_LIB_PATH; else export SGM_LIB_PATH=$P/libsgmpf_$v.so; fi
  echo "== $v rep $rep"
  timeout 600 python scripts/probe_models.py --dtype f32 --pf nemeth,filter 2>&1 | grep -E "nemeth|filter" | cut -c1-112
done
done
