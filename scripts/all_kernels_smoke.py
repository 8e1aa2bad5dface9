"""Small ragged cases of every kernel family x model x resampling scheme (fused single-CTA path, multi-CTA ragged tiles,
split-J, PaRIS queues, predictive statistic): everything must run and stay finite.  Also the target for compute-sanitizer
where that is available (it is closed on the build pool)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
rs = np.random.RandomState(0)
TH = {"svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0, 0, 0, 0, 0, 0, 0.5, 0.5], "prior"),
      "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0, 0, 0, 0, 0, 0.1, 1.0], "optimal"),
      "garch": ([0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09], "optimal")}
for model, (th, kern) in TH.items():
    it = sg.PFItems()
    for b in range(3):
        T = 7 + b
        it.add(rs.normal(size=T) * 0.7, th, t1=1, tL=T - 1, weights=1 + rs.rand(T - 2), prior_mean=0.0, prior_var=1.0)
    for N in (300, 2500, 5000):                  # fused single-CTA path / multi-CTA ragged / 3 CTAs per item
        for pf, kw in (("poyiadjis_N", {}), ("nemeth", dict(lambduh=0.9)), ("filter", {}), ("poyiadjis_N2", {}),
                       ("poyiadjis_N2", dict(n2_mode="fp32_pipe")), ("paris", dict(Ntilde=3))):
            for resample in ("multinomial_sorted", "multinomial", "systematic", "stratified"):
                r = sg.run_pf(model, kern, pf, it, N, dtype="f32", resample=resample, seed=1, offset=1, **kw)
                assert np.all(np.isfinite(r.grad)), (model, pf, N, resample)
        for pf, kw in (("poyiadjis_N", {}), ("nemeth", dict(lambduh=0.9)), ("filter", {})):      # f64 fast modes (ragged)
            for variates in ("native", "f32"):
                r = sg.run_pf(model, kern, pf, it, N, dtype="f64", variates=variates, seed=1, offset=1, **kw)
                assert np.all(np.isfinite(r.grad)), (model, pf, N, variates)
        r = sg.run_pf(model, kern, "poyiadjis_N", it, N, dtype="f64", seed=1, offset=1, want=("x", "lw", "stats", "anc", "trace_x"))
        r = sg.run_pf(model, kern, "filter", it, N, dtype="f32", stat_kind="pred", num_steps_ahead=4, seed=1, offset=1)
        assert np.all(np.isfinite(r.grad))
print("all kernels smoke ok")
