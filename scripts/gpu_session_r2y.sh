#!/bin/bash
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r2y_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2y_tests.log
tail -4 $O/r2y_tests.log
timeout 600 python scripts/probe_models.py --dtype f64 --variates native --json $O/r2y_models_f64.json > $O/r2y_models_f64.log 2>&1; cat $O/r2y_models_f64.log | tail -12
timeout 600 python scripts/probe_models.py --dtype f64 --variates f32 --pf poyiadjis_N --json $O/r2y_models_f64v32.json > $O/r2y_models_f64v32.log 2>&1; cat $O/r2y_models_f64v32.log | tail -4
