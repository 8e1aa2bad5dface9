"""Cluster kernel A/B (SGM_LIB_PATH selects the build): gradient latency with path='cluster' vs 'small' / 'tiles'."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
print("lib", os.environ.get("SGM_LIB_PATH", "default"))
ref = {}
for N in (1000, 1500, 2048, 3000, 4096, 8192, 16384, 65536):
    for path in ("cluster", "small", "tiles"):
        if (path == "small" and N > 2048) or (path == "cluster" and N > 16384):
            continue
        it = sg.PFItems()
        it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
        p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32", path=path, seed=3).upload()
        for k in range(3):
            p.launch(offset=k + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(20):
            p.launch(offset=10 + k)
        e1.record(); torch.cuda.synchronize()
        r = p.download().wait()
        print("N=%d %-8s %.4f ms  grad %s ll %.4f" % (N, path, e0.elapsed_time(e1) / 20, np.round(r.grad[0], 3), r.loglik[0]), flush=True)
