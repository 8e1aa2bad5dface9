"""Per-phase cycle counts of the cooperative kernel (needs a -DSGM_COOP_TIMING=1 build selected with SGM_LIB_PATH)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
for N in (8192, 65536):
    it = sg.PFItems()
    it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
    p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32").upload()
    print("N", N, flush=True)
    for k in range(3):
        p.launch(offset=k + 1)
        torch.cuda.synchronize()
