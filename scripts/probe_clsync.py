"""Cluster-synchronised single launch of the tile kernels against the grid-synchronised one / per-step launches:
gradient latency at 2048 < N <= 16384 for several batch sizes.  Run with SGM_NO_CLSYNC=1 for the other arm."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
print("SGM_NO_CLSYNC", os.environ.get("SGM_NO_CLSYNC"))
for N in (4096, 8192, 10000, 16384):
    for B in (1, 8, 25, 64, 128):
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
