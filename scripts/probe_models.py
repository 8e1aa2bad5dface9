"""O(N) step throughput of the three models (sorted multinomial, N = 2^16, 60 steps) with their algorithmic bytes per
particle-step (e (2 (n + p) + 2): SVM 40, LGSSM 48, GARCH 56 in f32, doubled in f64).

    python scripts/probe_models.py [--dtype f32|f64] [--variates native|f32] [--pf poyiadjis_N,nemeth,filter]
                                   [--models svm,lgssm,garch] [--items 256] [--particles 65536] [--json out.json]
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f32")
ap.add_argument("--variates", default="native")
ap.add_argument("--pf", default="poyiadjis_N,nemeth,filter")
ap.add_argument("--models", default="svm,lgssm,garch")
ap.add_argument("--items", type=int, default=256)
ap.add_argument("--particles", type=int, default=65536)
ap.add_argument("--json", default=None)
args = ap.parse_args()
rs = np.random.RandomState(0)
CASES = {"svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], "prior", 40, 10.0),
         "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0], "optimal", 48, 10.0),
         "garch": ([0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09], "optimal", 56, 1.0)}
B, N, T = args.items, args.particles, 60
peak = 6450.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = []
for model in args.models.split(","):
    th, kern, nbytes, pv = CASES[model]
    nbytes = nbytes * (2 if args.dtype == "f64" else 1)
    for pf in args.pf.split(","):
        kw = dict(lambduh=0.95) if pf == "nemeth" else {}
        it = sg.PFItems()
        for b in range(B):
            it.add(rs.normal(size=T) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=pv)
        prep = sg.engine.PreparedPF(model, kern, pf, it, N, dtype=args.dtype, rng="philox", resample="multinomial_sorted",
                                    variates=args.variates, **kw).upload()
        for k in range(2):
            prep.launch(offset=k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(3):
            prep.launch(offset=5 + k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ps = B * N * T / (ms * 1e-3)
        res = prep.download().wait()
        ok = bool(np.all(np.isfinite(res.grad)) and np.all(res.status == 0))
        rows.append(dict(model=model, pf=pf, dtype=args.dtype, variates=args.variates, items=B, particles=N, ms=ms,
                         particle_steps_per_s=ps, alg_bytes=nbytes, frac=ps * nbytes / (peak * 1e9), finite=ok,
                         lib=os.environ.get("SGM_LIB_PATH", "default")))
        print("%-6s %-12s %s/%s %8.3f ms  %.3e particle-steps/s  %d B/particle-step -> %.0f GB/s = %.1f %% of %.1f  finite=%s" % (
            model, pf, args.dtype, args.variates, ms, ps, nbytes, ps * nbytes / 1e9, 100 * ps * nbytes / (peak * 1e9), peak, ok), flush=True)
if args.json:
    with open(args.json, "w") as f:
        json.dump(rows, f, indent=1)
