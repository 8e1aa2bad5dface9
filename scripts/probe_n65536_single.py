"""One item, N = 2^16 (and 40000): which single-launch form runs and how long it takes (SGM_NO_CLSYNC=1: grid-barrier form)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
for N in (40000, 65536):
    it = sg.PFItems()
    it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
    p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32").upload()
    for k in range(3):
        p.launch(offset=k + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(10):
        r = p.launch(offset=10 + k)
    e1.record(); torch.cuda.synchronize()
    print("N=%d  %.4f ms  launches %s  SGM_NO_CLSYNC=%s" % (N, e0.elapsed_time(e1) / 10, getattr(r, "launches", "?"), os.environ.get("SGM_NO_CLSYNC")), flush=True)
