"""Single-launch kernel timing by batch size (N <= 2048), SVM poyiadjis_N f32; companion of probe_fused_crossover.py."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
th, T = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], 60
for N in (256, 1000, 1024, 2048):
    for B in (1, 148, 296, 444, 512, 768, 1024):
        it = sg.PFItems()
        for b in range(B):
            it.add(rs.normal(size=T) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=10.0)
        prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32", rng="philox", resample="multinomial_sorted").upload()
        for k in range(3):
            prep.launch(offset=k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(5):
            prep.launch(offset=5 + k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("N=%-5d B=%-5d %8.3f ms  %.3e p-s/s  launches %d" % (N, B, ms, B * N * T / (ms * 1e-3), prep.launches), flush=True)
