"""Quick throughput probe (not the bench): SVM poyiadjis_N, T_buf=60, various N / batch sizes."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg

rs = np.random.RandomState(0)
T = 10000
x = np.zeros(T); y = np.zeros(T); xp = 0.0
for t in range(T):
    xp = 0.95 * xp + np.sqrt(0.5) * rs.normal(); x[t] = xp; y[t] = np.sqrt(0.5) * np.exp(xp / 2) * rs.normal()
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]

def items(B):
    it = sg.PFItems()
    for b in range(B):
        s = rs.randint(10, T - 70)
        it.add(y[s - 10:s + 50], th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=10.0)
    return it

def run(model, pf, N, B, dtype, resample, reps=3, **kw):
    it = items(B)
    for _ in range(2):
        sg.run_pf(model, "prior", pf, it, N, dtype=dtype, resample=resample, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = sg.run_pf(model, "prior", pf, it, N, dtype=dtype, resample=resample, sync=False, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ps = N * 60 * B / (ms * 1e-3)
    print("%-6s %-12s N=%-8d B=%-5d %s %-18s %8.3f ms  %.3e particle-steps/s  (%.1f%% of 1.61e11)" % (
        model, pf, N, B, dtype, resample, ms, ps, 100 * ps * 40 / 6450.6e9), flush=True)

for N, B in [(65536, 1), (65536, 16), (65536, 128), (65536, 512), (1024, 1), (1024, 1024), (1 << 20, 8)]:
    for resample in ("multinomial_sorted", "multinomial"):
        run("svm", "poyiadjis_N", N, B, "f32", resample)
run("svm", "poyiadjis_N", 65536, 128, "f64", "multinomial_sorted")
run("svm", "poyiadjis_N", 65536, 128, "f32", "systematic")
run("svm", "poyiadjis_N2", 8192, 4, "f32", "multinomial_sorted", reps=1)
run("svm", "paris", 16384, 16, "f32", "multinomial_sorted", reps=1)
