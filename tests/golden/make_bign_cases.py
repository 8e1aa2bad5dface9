#!/usr/bin/env python3
"""Generate tests/golden/bign_cases.npz: outputs of the UNMODIFIED reference (imported read-only from /root/reference)
for the O(N^2) Poyiadjis smoother at N = 1024 / 2048 and the PaRIS smoother at N = 1024 / 4096 -- the sizes at which the
CUDA back ends leave their single-tile / single-split regime (multi-tile parent streaming, split-J grid, tensor-core
kernel; PaRIS round structure with thousands of entries) -- under fixed legacy seeds, in the layout of the `k/` section
of ref_cases.npz (same keys, so tests replay them with the same code).  Short windows keep the reference affordable
(pf.py:84-136 is O(N^2) Python: ~0.5 s per step at N = 1024, ~2 s at N = 2048).  Build-container only."""
import os
import sys
import time
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
warnings.filterwarnings("ignore")

from make_ref_cases import MODELS, theta_of  # noqa: E402  (parameter sets / generators of the k/ section)
from sgmcmc_ssm.particle_filters.buffered_smoother import buffered_pf_wrapper  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bign_cases.npz")
store = {}

CASES = [  # (model, kernel, pf, N, T_buf, extra)
    ("svm", "prior", "poyiadjis_N2", 1024, 7, {}),
    ("lgssm", "optimal", "poyiadjis_N2", 1024, 6, {}),
    ("garch", "optimal", "poyiadjis_N2", 1024, 6, {}),
    ("svm", "prior", "poyiadjis_N2", 2048, 5, {}),
    ("garch", "prior", "poyiadjis_N2", 2048, 5, {}),
    ("garch", "optimal", "paris", 1024, 12, {}),
    ("svm", "prior", "paris", 1024, 12, {}),
    ("lgssm", "prior", "paris", 1024, 10, {"Ntilde": 3}),
    ("garch", "optimal", "paris", 4096, 10, {}),
    ("svm", "prior", "paris", 4096, 8, {}),
]


def main():
    np.random.seed(4048)
    data = {m: MODELS[m]["gen"](T=200, parameters=MODELS[m]["params"]()) for m in MODELS}
    seed = 7000
    for model, kern, pf, N, Tb, extra in CASES:
        spec = MODELS[model]
        params = spec["params"]()
        helper = spec["Helper"](n=1, m=1)
        seed += 1
        name = "k/{0}_{1}_{2}_{3}_{4}".format(model, kern, pf, N, "_".join("%s%s" % kv for kv in extra.items()) or "d")
        t1, tL = 1, Tb - 1
        start = 23 + (seed % 40)
        window = data[model]["observations"][start:start + Tb]
        weights = np.linspace(0.5, 2.0, tL - t1)
        if model == "garch":
            prior_mean, prior_var = 0.0, float(params.alpha / (1 - params.beta - params.gamma))
        else:
            prior_mean, prior_var = 0.3, 1.7
        t0 = time.time()
        np.random.seed(seed)
        out = buffered_pf_wrapper(pf=pf, observations=window, parameters=params, N=N, kernel=helper._get_kernel(kern),
                                  additive_statistic_func=spec["score"], statistic_dim=spec["p"], t1=t1, tL=tL,
                                  weights=weights, prior_mean=prior_mean, prior_var=prior_var, **extra)
        for k, v in dict(seed=seed, obs=window, theta=theta_of(model, params), N=N, t1=t1, tL=tL, weights=weights,
                         prior_mean=prior_mean, prior_var=prior_var, x_t=out["x_t"], log_weights=out["log_weights"],
                         statistics=out["statistics"], loglik=out["loglikelihood_estimate"]).items():
            store[name + "/" + k] = np.asarray(v)
        for k, v in extra.items():
            store[name + "/opt_" + k] = np.asarray(v)
        print(name, float(out["loglikelihood_estimate"]), round(time.time() - t0, 1), "s", flush=True)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
