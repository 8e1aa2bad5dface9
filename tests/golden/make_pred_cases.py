#!/usr/bin/env python3
"""Generate tests/golden/pred_cases.npz: the UNMODIFIED reference's k-step-ahead predictive log-likelihood
(Helper.pf_predictive_loglikelihood_estimate, svm/helper.py:187-247, lgssm/helper.py:1048-1087,
garch/helper.py pf_predictive_loglikelihood_estimate) and the sampler-level
predictive_loglikelihood(kind='pf') (sgmcmc_sampler.py:94-126) under fixed legacy seeds.
Build-container only (imports /root/reference read-only); the .npz is committed."""
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")
import make_ref_cases as M  # noqa: E402  (re-uses its parameter builders)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pred_cases.npz")
store = {}
SAMPLERS = {"svm": M.SVMSampler, "lgssm": M.LGSSMSampler, "garch": M.GARCHSampler}
for model, spec in M.MODELS.items():
    for tag, (T, t1, tL, K, N, seed) in {"a": (30, 5, 25, 3, 300, 11), "b": (12, 0, 12, 5, 64, 12), "c": (20, 4, 20, 5, 1000, 13),
                                            "d": (40, 5, 35, 10, 200, 14)}.items():   # d: the default horizon (10)
        np.random.seed(1000 + seed)
        p = spec["params"]()
        data = spec["gen"](T=T + 5, parameters=p)
        obs = data["observations"][:T]
        h = spec["Helper"](**p.dim)
        for kernel in spec["kernels"]:
            np.random.seed(seed)
            out = h.pf_predictive_loglikelihood_estimate(obs, p, num_steps_ahead=K, subsequence_start=t1,
                                                         subsequence_end=tL, N=N, kernel=kernel)
            pre = "p/{0}_{1}_{2}".format(model, kernel, tag)
            for k, v in dict(obs=obs, theta=M.theta_of(model, p), Q=float(np.ravel(p.Q)[0]) if model != "garch" else 0.0,
                             R=float(np.ravel(p.R)[0]), t1=t1, tL=tL, K=K, N=N, seed=seed, out=out).items():
                store[pre + "/" + k] = np.asarray(v)
    # sampler level
    np.random.seed(77)
    p = spec["params"]()
    data = spec["gen"](T=200, parameters=p)
    s = SAMPLERS[model](n=1, m=1, observations=data["observations"], parameters=p.copy()) if model != "garch" else \
        SAMPLERS[model](n=1, m=1, observations=data["observations"], parameters=p.copy())
    np.random.seed(5)
    out = s.predictive_loglikelihood(kind="pf", num_steps_ahead=4, subsequence_length=20, minibatch_size=3, buffer_length=4, N=200)
    pre = "ps/{0}".format(model)
    for k, v in dict(obs=data["observations"], theta=M.theta_of(model, p), out=out).items():
        store[pre + "/" + k] = np.asarray(v)
    np.random.seed(6)                         # default num_steps_ahead (10), as the drivers' metric gets it (lag= is ignored)
    out = s.predictive_loglikelihood(kind="pf", subsequence_length=30, minibatch_size=2, buffer_length=4, N=100)
    store["ps10/{0}/out".format(model)] = np.asarray(out)
np.savez_compressed(OUT, **store)
print("wrote", OUT, len(store), "arrays")
