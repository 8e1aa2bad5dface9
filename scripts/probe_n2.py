"""Throughput probe of the O(N^2) smoother (pair-steps/s) and PaRIS (particle-steps/s)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
TH = {"svm": [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0],
      "lgssm": [0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0],
      "garch": [0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09]}

def run(model, pf, N, B, T, reps=2, dtype="f32", hetero=False, **kw):
    y = rs.normal(size=T) * 0.7
    it = sg.PFItems()
    for b in range(B):
        if hetero:                       # every item its own series (what a minibatch of subsequences looks like)
            y = rs.normal(size=T) * 0.7
        it.add(y, TH[model], t1=2, tL=T - 2, weights=np.ones(T - 4) * 250.0, prior_mean=0.0, prior_var=1.0 if model == "garch" else 10.0)
    kern = "prior" if model == "svm" else "optimal"
    r = sg.run_pf(model, kern, pf, it, N, dtype=dtype, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = sg.run_pf(model, kern, pf, it, N, dtype=dtype, sync=False, **kw)
    e1.record(); torch.cuda.synchronize(); r.wait(check=False)
    ms = e0.elapsed_time(e1) / reps
    ps = N * T * B / (ms * 1e-3)
    extra = "  %.3e pair-steps/s" % (ps * N) if pf == "poyiadjis_N2" else ""
    print("%-6s %-12s N=%-7d B=%-4d T=%-3d %s %9.3f ms  %.3e particle-steps/s%s  status=%s grad0=%s" % (
        model, pf, N, B, T, dtype, ms, ps, extra, int(np.max(r.status)), np.array2string(r.grad[0][:4], precision=4)), flush=True)

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "one":        # one <model> <pf> <N> <B> <T> [n2_mode]   (ncu target)
        kw = dict(n2_mode=sys.argv[7]) if len(sys.argv) > 7 and sys.argv[7] != "hetero" else {}
        if "hetero" in sys.argv:
            kw["hetero"] = True
        run(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), reps=1, **kw)
    if what in ("n2", "all"):
        run("svm", "poyiadjis_N2", 1024, 1, 20)
        run("svm", "poyiadjis_N2", 16384, 8, 10, n2_mode="fp32_pipe")
        run("svm", "poyiadjis_N2", 8192, 4, 20)
        run("svm", "poyiadjis_N2", 16384, 8, 10)
        run("svm", "poyiadjis_N2", 65536, 1, 6)
        run("svm", "poyiadjis_N2", 65536, 4, 6)
        run("lgssm", "poyiadjis_N2", 16384, 8, 10)
        run("garch", "poyiadjis_N2", 16384, 8, 10)
        run("svm", "poyiadjis_N", 8192, 4, 20)
    if what in ("paris", "all"):
        run("garch", "paris", 16384, 1, 20)
        run("garch", "paris", 16384, 16, 20)
        run("garch", "paris", 16384, 128, 20)
        run("svm", "paris", 16384, 16, 20)
        run("svm", "paris", 65536, 16, 20)
