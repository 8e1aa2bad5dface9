"""ncu targets of round 2:  python scripts/profile_target2.py small|f64|f64v32|sgld"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
what = sys.argv[1]
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
def items(B, T=60):
    it = sg.PFItems()
    for b in range(B):
        it.add(rs.normal(size=T) * 0.7, th, t1=2, tL=T - 2, weights=np.ones(T - 4) * 25.0, prior_mean=0.0, prior_var=10.0)
    return it
if what == "small":
    p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", items(1), 1000, dtype="f32").upload()
    for k in range(4):
        p.launch(offset=k + 1)
elif what == "smallbatch":
    p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", items(2048), 1000, dtype="f32").upload()
    for k in range(4):
        p.launch(offset=k + 1)
elif what in ("f64", "f64v32"):
    p = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", items(256, 12), 65536, dtype="f64",
                             variates="f32" if what == "f64v32" else "native").upload()
    for k in range(2):
        p.launch(offset=k + 1)
elif what == "paris":
    gth = [0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09]
    it = sg.PFItems()
    for b in range(64):
        it.add(rs.normal(size=8) * 0.7, gth, t1=1, tL=7, prior_mean=0.0, prior_var=1.0)
    p = sg.engine.PreparedPF("garch", "optimal", "paris", it, 1 << 14, dtype="f32").upload()
    for k in range(2):
        p.launch(offset=k + 1)
torch.cuda.synchronize()
print("done", what)
