"""Gradient / filter latency of the GENERIC instantiation of the shared-memory kernel (Nemeth with shrinkage, filter, systematic
resampling) at N = 512 / 1000, 1-296 items (SGM_LIB_PATH selects the build)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0]
print("lib", os.environ.get("SGM_LIB_PATH", "default"))
for pf, kw in (("nemeth", dict(lambduh=0.95)), ("filter", {}), ("poyiadjis_N", dict(resample="systematic"))):
    for N in (512, 1000):
        for B in (1, 64, 296):
            it = sg.PFItems()
            for b in range(B):
                it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
            p = sg.engine.PreparedPF("lgssm", "optimal", pf, it, N, dtype="f32", path="small", **kw).upload()
            for k in range(3):
                p.launch(offset=k + 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for k in range(20):
                p.launch(offset=10 + k)
            e1.record(); torch.cuda.synchronize()
            print("%-12s N=%d B=%d  %.4f ms" % (pf, N, B, e0.elapsed_time(e1) / 20), flush=True)
