"""Small fixed workload for ncu: SVM poyiadjis_N, N=65536, B items, T_buf=60, f32 Philox."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
resample = sys.argv[3] if len(sys.argv) > 3 else "multinomial_sorted"
pf = sys.argv[4] if len(sys.argv) > 4 else "poyiadjis_N"
model = sys.argv[5] if len(sys.argv) > 5 else "svm"
rs = np.random.RandomState(0)
y = rs.normal(size=60) * 0.7
th = {"svm": [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0],
      "lgssm": [0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0],
      "garch": [0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09]}[model]
it = sg.PFItems()
for b in range(B):
    it.add(y, th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=1.0 if model == "garch" else 10.0)
kern = "prior" if model == "svm" else "optimal"
for _ in range(2):
    r = sg.run_pf(model, kern, pf, it, N, dtype="f32", resample=resample, seed=1)
torch.cuda.synchronize()
print("ok", r.grad[0], r.launches)
