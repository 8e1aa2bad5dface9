#!/usr/bin/env python3
"""Generate tests/golden/model_sweep_stats.npz: the distribution (mean / std over independent repetitions) of the
UNMODIFIED reference's buffered particle-filter gradient estimator, for the model / kernel / smoother combinations
whose production (f32, Philox, order-statistics resampling) instantiations have no stored reference sweep to be
pinned to (the reference's scratch/ holds SVM poyiadjis_N only; see make_svm_sweep_stats.py for that one):

    LGSSM prior / optimal, GARCH prior / optimal   x poyiadjis_N            (protocol of *_grad_compare.py:79-129:
    SVM prior, LGSSM optimal, GARCH optimal        x nemeth (lambduh 0.95)   L = 16 subsequence in the middle of a
    SVM prior, LGSSM optimal, GARCH optimal        x filter                  T = 100 series, buffers B on both sides,
    GARCH optimal, SVM prior                       x paris (Ntilde = 2)      REPS repetitions per (B, N) cell)

N covers a ragged particle count (1000, 10000: not multiples of the 256-particle warp tile) and a full-tile one (4096).
Every cell is reference output: `sgmcmc_ssm.particle_filters.buffered_smoother.buffered_pf_wrapper` +
`average_statistic` (buffered_smoother.py:12-199) with the model's own kernel and complete-data score function, under
`np.random.seed(cell seed + repetition)`.  Build-container only (imports /root/reference); the .npz is committed.
"""
import multiprocessing as mp
import os
import sys
import time
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from sgmcmc_ssm.models.svm import SVMParameters, SVMHelper, generate_svm_data  # noqa: E402
from sgmcmc_ssm.models.lgssm import LGSSMParameters, LGSSMHelper, generate_lgssm_data  # noqa: E402
from sgmcmc_ssm.models.garch import GARCHParameters, GARCHHelper, generate_garch_data  # noqa: E402
from sgmcmc_ssm.models.svm.helper import svm_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.models.lgssm.helper import lgssm_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.models.garch.helper import garch_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.particle_filters.buffered_smoother import buffered_pf_wrapper, average_statistic  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_sweep_stats.npz")
L, T = 16, 100


def svm_params():
    return SVMParameters(A=np.eye(1) * 0.95, LQinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.5)),
                         LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.5)))


def lgssm_params():
    return LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.1)),
                           LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 1.0)))


def garch_params():
    log_mu, logit_phi, logit_lambduh = GARCHParameters.convert_alpha_beta_gamma(0.1, 0.8, 0.05)
    return GARCHParameters(log_mu=log_mu, logit_phi=logit_phi, logit_lambduh=logit_lambduh,
                           LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.3 ** 2)))


def theta_of(model, p):
    f = lambda a: float(np.ravel(a)[0])  # noqa: E731
    if model == "svm":
        return [f(p.A), f(p.LQinv), f(p.Qinv), f(p.LRinv), f(p.Rinv)]
    if model == "lgssm":
        return [f(p.A), f(p.LQinv), f(p.Qinv), f(p.C), f(p.LRinv), f(p.Rinv)]
    return [f(p.alpha), f(p.beta), f(p.gamma), f(p.mu), f(p.phi), f(p.lambduh), f(p.LRinv), f(p.Rinv), f(p.R)]


MODELS = {
    "svm": dict(params=svm_params, gen=generate_svm_data, Helper=SVMHelper, score=svm_complete_data_loglike_gradient, p=3),
    "lgssm": dict(params=lgssm_params, gen=generate_lgssm_data, Helper=LGSSMHelper,
                  score=lgssm_complete_data_loglike_gradient, p=4),
    "garch": dict(params=garch_params, gen=generate_garch_data, Helper=GARCHHelper,
                  score=garch_complete_data_loglike_gradient, p=4),
}

# (model, kernel, pf, extra kwargs, Ns, buffer sizes, repetitions)
GROUPS = [
    ("lgssm", "prior", "poyiadjis_N", {}, [1000, 4096, 10000], [10, 4, 0], 64),
    ("lgssm", "optimal", "poyiadjis_N", {}, [1000, 4096, 10000], [10, 4, 0], 64),
    ("garch", "prior", "poyiadjis_N", {}, [1000, 4096, 10000], [10, 4, 0], 64),
    ("garch", "optimal", "poyiadjis_N", {}, [1000, 4096, 10000], [10, 4, 0], 64),
    ("svm", "prior", "nemeth", {}, [1000, 4096], [10, 0], 64),
    ("lgssm", "optimal", "nemeth", {}, [1000, 4096], [10, 0], 64),
    ("garch", "optimal", "nemeth", {}, [1000, 4096], [10, 0], 64),
    ("svm", "prior", "filter", {}, [1000, 4096], [10, 0], 64),
    ("lgssm", "optimal", "filter", {}, [1000, 4096], [10, 0], 64),
    ("garch", "optimal", "filter", {}, [1000, 4096], [10, 0], 64),
    ("garch", "optimal", "paris", {}, [1000, 4096], [8, 0], 48),
    ("svm", "prior", "paris", {}, [1000, 4096], [8, 0], 48),
]

DATA = {}


def make_data():
    np.random.seed(31337)
    for m, spec in MODELS.items():
        DATA[m] = spec["gen"](T=T, parameters=spec["params"]())["observations"]


def prior_moments(model, params):
    if model == "garch":        # garch/helper.py:324-332 (_get_prior_x with forward_message None)
        return 0.0, float(params.alpha / (1 - params.beta - params.gamma))
    return 0.0, 10.0            # default forward message: precision I / 10 (svm/helper.py:31-36, lgssm/helper.py)


def one(task):
    gi, N, B, rep, seed = task
    model, kern, pf, extra, _, _, _ = GROUPS[gi]
    spec = MODELS[model]
    params = spec["params"]()
    helper = spec["Helper"](n=1, m=1)
    obs = DATA[model]
    t0 = (T + L) // 2
    window = obs[t0 - B:t0 + L + B]
    pm, pv = prior_moments(model, params)
    np.random.seed(seed)
    out = buffered_pf_wrapper(pf=pf, observations=window, parameters=params, N=N, kernel=helper._get_kernel(kern),
                              additive_statistic_func=spec["score"], statistic_dim=spec["p"], t1=B, tL=L + B,
                              weights=None, prior_mean=pm, prior_var=pv, **extra)
    if pf == "filter":
        g = np.asarray(out["statistics"], dtype=float)
    else:
        g = average_statistic(out)
    return gi, N, B, rep, np.asarray(g, dtype=float), float(out["loglikelihood_estimate"])


def main():
    make_data()
    tasks, seed = [], 100000
    for gi, (model, kern, pf, extra, Ns, Bs, reps) in enumerate(GROUPS):
        for N in Ns:
            for B in Bs:
                for rep in range(reps):
                    seed += 1
                    tasks.append((gi, N, B, rep, seed))
    # expensive cells first so the pool stays busy until the end
    cost = lambda t: (GROUPS[t[0]][2] == "paris") * 50 * t[1] + t[1]  # noqa: E731
    tasks.sort(key=cost, reverse=True)
    t0 = time.time()
    with mp.get_context("fork").Pool(os.cpu_count() or 1) as pool:
        results = pool.map(one, tasks, chunksize=2)
    print("ran", len(tasks), "reference gradients in", round(time.time() - t0, 1), "s", flush=True)
    store = {}
    for gi, (model, kern, pf, extra, Ns, Bs, reps) in enumerate(GROUPS):
        name = "{0}_{1}_{2}".format(model, kern, pf)
        p = MODELS[model]["p"]
        g = np.zeros((len(Ns), len(Bs), reps, p))
        ll = np.zeros((len(Ns), len(Bs), reps))
        for r in results:
            if r[0] == gi:
                g[Ns.index(r[1]), Bs.index(r[2]), r[3]] = r[4]
                ll[Ns.index(r[1]), Bs.index(r[2]), r[3]] = r[5]
        params = MODELS[model]["params"]()
        pm, pv = prior_moments(model, params)
        store[name + "/Ns"] = np.array(Ns)
        store[name + "/buffer_sizes"] = np.array(Bs)
        store[name + "/reps"] = np.array(reps)
        store[name + "/mean"] = g.mean(axis=2)
        store[name + "/std"] = g.std(axis=2, ddof=1)
        store[name + "/loglik_mean"] = ll.mean(axis=2)
        store[name + "/loglik_std"] = ll.std(axis=2, ddof=1)
        store[name + "/theta"] = np.array(theta_of(model, params))
        store[name + "/prior_mean"] = np.array(pm)
        store[name + "/prior_var"] = np.array(pv)
        store[name + "/obs"] = DATA[model]
        print(name, g.mean(axis=2)[0, 0], g.std(axis=2, ddof=1)[0, 0], flush=True)
    store["L"] = np.array(L)
    store["t0"] = np.array((T + L) // 2)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
