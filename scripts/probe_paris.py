"""PaRIS (device randoms) diagnostics: time per gradient, accept-reject proposals per (particle, replicate, step) and entries
resolved by the exact sampler, from the counters the kernels keep (PFResult.diag)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(1)
CASES = {"garch": ([0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09], "optimal", 1.0),
         "svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], "prior", 10.0),
         "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0], "optimal", 10.0)}
out = {}
for model, (th, kern, pv) in CASES.items():
    for N, B, T in ((1 << 14, 64, 60), (1 << 12, 64, 60)):
        it = sg.PFItems()
        for _ in range(B):
            it.add(rs.normal(size=T) * 0.7, th, t1=2, tL=T - 2, prior_mean=0.0, prior_var=pv)
        p = sg.engine.PreparedPF(model, kern, "paris", it, N, dtype="f32", rng="philox").upload()
        p.launch(offset=1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(2):
            p.launch(offset=2 + k)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        r = p.download().wait()
        entries = float(N) * 2 * T
        key = "%s_N%d_B%d" % (model, N, B)
        out[key] = dict(ms_per_gradient=ms, particle_steps_per_s=B * N * T / (ms * 1e-3),
                        proposals_per_entry=float(r.diag[:, 0].mean() / entries),
                        exact_fraction=float(r.diag[:, 1].mean() / entries), launches=p.launches)
        print(key, out[key], flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
