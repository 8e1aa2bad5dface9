#!/bin/bash
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2i_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2i_tests.log
tail -4 $O/r2i_tests.log
python scripts/probe_paris.py $O/r2i_paris.json 2>&1 | tail -8
python scripts/probe_models.py --dtype f64 --models svm,lgssm,garch --pf poyiadjis_N --json $O/r2i_probe_f64.json 2>&1 | tail -4
