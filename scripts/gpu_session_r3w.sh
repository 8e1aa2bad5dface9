#!/bin/bash
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize_properties.py tests/test_gpu_edge_cases.py -m gpu -q -p no:cacheprovider > $O/r3w_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3w_tests.log
tail -4 $O/r3w_tests.log
python - <<'PY'
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
for N in (16384, 20000, 24576, 32768, 40000, 65536):
    for B in (1, 4, 9):
        it = sg.PFItems()
        for b in range(B):
            it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
        p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32").upload()
        for k in range(3):
            p.launch(offset=k + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(10):
            r = p.launch(offset=10 + k)
        e1.record(); torch.cuda.synchronize()
        print("N=%d B=%d  %.4f ms  launches %s" % (N, B, e0.elapsed_time(e1) / 10, getattr(r, "launches", "?")), flush=True)
PY
