"""O(N) step throughput of the three models (f32, sorted multinomial, N = 2^16, 256 items, 60 steps) with their
algorithmic bytes per particle-step (4 (2 (n + p) + 2): SVM 40, LGSSM 48, GARCH 56)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
CASES = {"svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], "prior", 40, 10.0),
         "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0], "optimal", 48, 10.0),
         "garch": ([0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09], "optimal", 56, 1.0)}
B, N, T = 256, 65536, 60
for model, (th, kern, nbytes, pv) in CASES.items():
    for pf, kw in (("poyiadjis_N", {}), ("nemeth", dict(lambduh=0.95)), ("filter", {})):
        it = sg.PFItems()
        for b in range(B):
            it.add(rs.normal(size=T) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=pv)
        prep = sg.engine.PreparedPF(model, kern, pf, it, N, dtype="f32", rng="philox", resample="multinomial_sorted", **kw).upload()
        for k in range(2):
            prep.launch(offset=k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(3):
            prep.launch(offset=5 + k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ps = B * N * T / (ms * 1e-3)
        print("%-6s %-12s %8.3f ms  %.3e particle-steps/s  %d B/particle-step -> %.0f GB/s = %.1f %% of 6450.6" % (
            model, pf, ms, ps, nbytes, ps * nbytes / 1e9, 100 * ps * nbytes / 6450.6e9), flush=True)
