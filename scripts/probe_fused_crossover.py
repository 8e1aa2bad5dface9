"""Fused single-launch kernel vs the per-step kernels for N <= 2048, by batch size (run once per library build:
SGM_LIB_PATH selects the build).  SVM / LGSSM, poyiadjis_N, f32, 60-step windows."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
CASES = {"svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], "prior"), "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0], "optimal")}
T = 60
for model in sys.argv[1:] or ["svm"]:
    th, kern = CASES[model]
    for N in (256, 512, 1000, 1024, 2048):
        for B in (1, 8, 32, 128, 512, 2048, 8192):
            it = sg.PFItems()
            for b in range(B):
                it.add(rs.normal(size=T) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=10.0)
            prep = sg.engine.PreparedPF(model, kern, "poyiadjis_N", it, N, dtype="f32", rng="philox", resample="multinomial_sorted").upload()
            for k in range(3):
                prep.launch(offset=k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for k in range(5):
                prep.launch(offset=5 + k)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("%s %s N=%-5d B=%-5d %8.3f ms  %.3e p-s/s" % (os.environ.get("SGM_LIB_PATH", "default")[-12:], model, N, B, ms, B * N * T / (ms * 1e-3)), flush=True)
