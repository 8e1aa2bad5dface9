#!/usr/bin/env python3
"""Generate tests/golden/ref_cases.npz by running the UNMODIFIED reference (imported read-only from
/root/reference) under fixed legacy seeds.  Build-container only; the .npz is committed.

Sections (keys are prefixed):
  k/<name>/...   kernel level: particle_filters.buffered_smoother.buffered_pf_wrapper for every
                 (model, kernel, smoother) combination  -> x_t, log_weights, statistics, loglik
  h/<name>/...   helper level: pf_gradient_estimate / pf_loglikelihood_estimate /
                 pf_latent_var_distr (default priors, default kernels)
  s/<name>/...   sampler level: noisy_gradient(kind='pf'), sample_sgld + project_parameters,
                 LGSSM sample_sgrld, SeqSVMSampler; prior logprior / grad_logprior;
                 random_subsequence_and_weights
  a/<name>/...   analytic LGSSM (Kalman) buffered gradient -- the known-answer for large-N PF

Each case stores its seed and every input needed to replay it without the reference.
"""
import os
import sys
import warnings

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

from sgmcmc_ssm.models.svm import (SVMParameters, SVMHelper, SVMSampler, SeqSVMSampler,  # noqa: E402
                                   SVMPrior, generate_svm_data)
from sgmcmc_ssm.models.lgssm import (LGSSMParameters, LGSSMHelper, LGSSMSampler,  # noqa: E402
                                     LGSSMPrior, generate_lgssm_data)
from sgmcmc_ssm.models.lgssm.parameters import LGSSMPreconditioner  # noqa: E402
from sgmcmc_ssm.models.garch import (GARCHParameters, GARCHHelper, GARCHSampler,  # noqa: E402
                                     GARCHPrior, generate_garch_data)
from sgmcmc_ssm.models.svm.helper import svm_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.models.lgssm.helper import lgssm_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.models.garch.helper import garch_complete_data_loglike_gradient  # noqa: E402
from sgmcmc_ssm.particle_filters.buffered_smoother import buffered_pf_wrapper  # noqa: E402
from sgmcmc_ssm.sgmcmc_sampler import random_subsequence_and_weights  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cases.npz")
store = {}


def put(prefix, **kv):
    for k, v in kv.items():
        store[prefix + "/" + k] = np.asarray(v)


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
    f = lambda a: float(np.ravel(a)[0])
    if model == "svm":
        return [f(p.A), f(p.LQinv), f(p.Qinv), f(p.LRinv), f(p.Rinv)]
    if model == "lgssm":
        return [f(p.A), f(p.LQinv), f(p.Qinv), f(p.C), f(p.LRinv), f(p.Rinv)]
    return [f(p.alpha), f(p.beta), f(p.gamma), f(p.mu), f(p.phi), f(p.lambduh), f(p.LRinv), f(p.Rinv), f(p.R)]


MODELS = {
    "svm": dict(params=svm_params, gen=generate_svm_data, Helper=SVMHelper,
                score=svm_complete_data_loglike_gradient, p=3, kernels=["prior"]),
    "lgssm": dict(params=lgssm_params, gen=generate_lgssm_data, Helper=LGSSMHelper,
                  score=lgssm_complete_data_loglike_gradient, p=4, kernels=["prior", "optimal"]),
    "garch": dict(params=garch_params, gen=generate_garch_data, Helper=GARCHHelper,
                  score=garch_complete_data_loglike_gradient, p=4, kernels=["prior", "optimal"]),
}


def kernel_level():
    np.random.seed(2024)
    data = {m: MODELS[m]["gen"](T=200, parameters=MODELS[m]["params"]()) for m in MODELS}
    seed = 1000
    for model, spec in MODELS.items():
        params = spec["params"]()
        obs = data[model]["observations"]
        helper = spec["Helper"](n=1, m=1)
        for kern in spec["kernels"]:
            for pf, N, Tb, extra in [
                ("poyiadjis_N", 64, 14, {}),
                ("poyiadjis_N", 1000, 24, {}),
                ("nemeth", 200, 14, {}),
                ("nemeth", 200, 14, {"lambduh": 0.8}),
                ("poyiadjis_N2", 48, 10, {}),
                ("paris", 300, 12, {}),
                ("paris", 300, 12, {"Ntilde": 3}),
                ("paris", 40, 8, {"accept_reject": False}),
                ("filter", 128, 12, {}),
            ]:
                seed += 1
                name = "k/{0}_{1}_{2}_{3}_{4}".format(model, kern, pf, N, "_".join(
                    "%s%s" % (k, v) for k, v in extra.items()) or "d")
                t1, tL = 3, Tb - 3
                start = 17 + (seed % 50)
                window = obs[start:start + Tb]
                weights = np.linspace(0.5, 2.0, tL - t1)
                if model == "garch":
                    prior_mean, prior_var = 0.0, float(params.alpha / (1 - params.beta - params.gamma))
                else:
                    prior_mean, prior_var = 0.3, 1.7
                np.random.seed(seed)
                out = buffered_pf_wrapper(
                    pf=pf, observations=window, parameters=params, N=N,
                    kernel=helper._get_kernel(kern), additive_statistic_func=spec["score"],
                    statistic_dim=spec["p"], t1=t1, tL=tL, weights=weights,
                    prior_mean=prior_mean, prior_var=prior_var, save_all=(N <= 64), **extra)
                put(name, seed=seed, obs=window, theta=theta_of(model, params), N=N, t1=t1, tL=tL,
                    weights=weights, prior_mean=prior_mean, prior_var=prior_var,
                    x_t=out["x_t"], log_weights=out["log_weights"], statistics=out["statistics"],
                    loglik=out["loglikelihood_estimate"])
                for k, v in extra.items():
                    put(name, **{"opt_" + k: v})
                if N <= 64:
                    put(name, all_x_t=out["all_x_t"], all_log_weights=out["all_log_weights"],
                        all_statistics=out["all_statistics"],
                        all_loglik=out["all_loglikelihood_estimate"])
                print(name, float(out["loglikelihood_estimate"]), flush=True)
    return data


def helper_level(data):
    seed = 5000
    for model, spec in MODELS.items():
        params = spec["params"]()
        obs = data[model]["observations"]
        for fm in ("default", "init"):
            forward_message = None if fm == "default" else data[model]["initial_message"]
            helper = spec["Helper"](n=1, m=1, forward_message=forward_message)
            for pf, N in [("poyiadjis_N", 500), ("paris", 200), ("nemeth", 200)]:
                seed += 1
                window = obs[40:70]
                weights = np.linspace(1.0, 3.0, 20)
                np.random.seed(seed)
                grad = helper.pf_gradient_estimate(observations=window, parameters=params,
                                                   subsequence_start=5, subsequence_end=25,
                                                   weights=weights, pf=pf, N=N,
                                                   tqdm=None, unknown_kwarg=3)
                name = "h/grad_{0}_{1}_{2}".format(model, fm, pf)
                put(name, seed=seed, obs=window, theta=theta_of(model, params), N=N, t1=5, tL=25,
                    weights=weights, keys=np.array(sorted(grad.keys())),
                    values=[float(np.ravel(grad[k])[0]) for k in sorted(grad.keys())])
                if forward_message is not None:
                    put(name, fm_precision=forward_message["precision"],
                        fm_mean_precision=forward_message["mean_precision"])
                print(name, flush=True)
            seed += 1
            np.random.seed(seed)
            ll = helper.pf_loglikelihood_estimate(observations=obs[40:70], parameters=params,
                                                  subsequence_start=5, subsequence_end=25,
                                                  weights=None, pf="poyiadjis_N", N=400)
            put("h/loglik_{0}_{1}".format(model, fm), seed=seed, obs=obs[40:70],
                theta=theta_of(model, params), N=400, t1=5, tL=25, value=float(ll))
            seed += 1
            np.random.seed(seed)
            mean, cov = helper.pf_latent_var_distr(observations=obs[40:60], parameters=params,
                                                   subsequence_start=4, subsequence_end=16,
                                                   pf="poyiadjis_N", N=300)
            put("h/latent_{0}_{1}".format(model, fm), seed=seed, obs=obs[40:60],
                theta=theta_of(model, params), N=300, t1=4, tL=16, mean=mean, cov=cov)
            if forward_message is not None:
                for nm in ("h/loglik_{0}_{1}".format(model, fm), "h/latent_{0}_{1}".format(model, fm)):
                    put(nm, fm_precision=forward_message["precision"],
                        fm_mean_precision=forward_message["mean_precision"])


def params_vec(p):
    return np.concatenate([np.ravel(p.var_dict[k]) for k in sorted(p.var_dict)])


def sampler_level(data):
    # subsequence weights
    for i, (S, T, seed) in enumerate([(10, 100, 1), (10, 100, 7), (40, 1000, 3), (16, 20, 5),
                                      (10, 100, 11), (10, 100, 12), (5, 8, 2)]):
        for style in ("uniform", "strict", "naive"):
            if style == "strict" and T % S != 0:
                continue
            np.random.seed(seed)
            s, e, w = random_subsequence_and_weights(S=S, T=T, partition_style=style)
            put("s/subseq_{0}_{1}".format(i, style), S=S, T=T, seed=seed, start=s, end=e, weights=w)

    specs = [("svm", SVMSampler, SVMPrior), ("lgssm", LGSSMSampler, LGSSMPrior),
             ("garch", GARCHSampler, GARCHPrior)]
    seed = 9000
    for model, Sampler, Prior in specs:
        params = MODELS[model]["params"]()
        obs = data[model]["observations"]
        sampler = Sampler(n=1, m=1, observations=obs, parameters=params.copy())
        prior = sampler.prior
        put("s/prior_" + model, **{"hyper_" + k: v for k, v in prior.hyperparams.items()})
        gl = prior.grad_logprior(parameters=params)
        put("s/prior_" + model, logprior=prior.logprior(params),
            grad_keys=np.array(sorted(gl)), grad_values=np.concatenate([np.ravel(gl[k]) for k in sorted(gl)]),
            param_keys=np.array(sorted(params.var_dict)), param_values=params_vec(params))
        for pf, N, mb in [("poyiadjis_N", 300, 2), ("nemeth", 150, 1)]:
            seed += 1
            sampler.parameters = params.copy()
            np.random.seed(seed)
            g = sampler.noisy_gradient(kind="pf", pf=pf, N=N, subsequence_length=12, buffer_length=4,
                                       minibatch_size=mb)
            name = "s/noisy_grad_{0}_{1}".format(model, pf)
            put(name, seed=seed, N=N, minibatch_size=mb, obs=obs, theta=theta_of(model, params),
                keys=np.array(sorted(g)), values=np.concatenate([np.ravel(g[k]) for k in sorted(g)]))
            print(name, flush=True)
        # SGLD step + projection
        seed += 1
        sampler.parameters = params.copy()
        np.random.seed(seed)
        for _ in range(3):
            sampler.sample_sgld(epsilon=0.01, kind="pf", pf="poyiadjis_N", N=200,
                                subsequence_length=12, buffer_length=4, minibatch_size=1)
            sampler.project_parameters()
        put("s/sgld_" + model, seed=seed, obs=obs, param_keys=np.array(sorted(params.var_dict)),
            before=params_vec(params), after=params_vec(sampler.parameters))
        print("s/sgld_" + model, params_vec(sampler.parameters), flush=True)
        if model == "lgssm":
            seed += 1
            sampler.parameters = params.copy()
            np.random.seed(seed)
            for _ in range(3):
                sampler.sample_sgrld(epsilon=0.01, preconditioner=LGSSMPreconditioner(), kind="pf",
                                     pf="poyiadjis_N", N=200, subsequence_length=12,
                                     buffer_length=4, minibatch_size=1)
                sampler.project_parameters()
            put("s/sgrld_lgssm", seed=seed, obs=obs, param_keys=np.array(sorted(params.var_dict)),
                before=params_vec(params), after=params_vec(sampler.parameters))
        # noisy loglikelihood (f1)
        seed += 1
        sampler.parameters = params.copy()
        np.random.seed(seed)
        ll = sampler.noisy_loglikelihood(kind="pf", pf="poyiadjis_N", N=250, subsequence_length=20,
                                         buffer_length=5, minibatch_size=2)
        put("s/noisy_loglik_" + model, seed=seed, obs=obs, N=250, value=float(ll))

    # Seq sampler (list of sequences)
    obs = data["svm"]["observations"]
    seqs = [obs[0:60], obs[60:110], obs[110:200]]
    params = svm_params()
    sampler = SeqSVMSampler(n=1, m=1, observations=seqs, parameters=params.copy())
    np.random.seed(777)
    g = sampler.noisy_gradient(kind="pf", pf="poyiadjis_N", N=200, subsequence_length=10,
                               buffer_length=3, minibatch_size=1, num_sequences=2)
    put("s/seq_svm", seed=777, N=200, lens=[60, 50, 90], obs=obs,
        keys=np.array(sorted(g)), values=np.concatenate([np.ravel(g[k]) for k in sorted(g)]))
    np.random.seed(778)
    g = sampler.noisy_gradient(kind="pf", pf="poyiadjis_N", N=200, subsequence_length=10,
                               buffer_length=3, minibatch_size=1)
    put("s/seq_svm_all", seed=778, N=200, lens=[60, 50, 90], obs=obs,
        keys=np.array(sorted(g)), values=np.concatenate([np.ravel(g[k]) for k in sorted(g)]))


def analytic_level(data):
    params = lgssm_params()
    obs = data["lgssm"]["observations"]
    sampler = LGSSMSampler(n=1, m=1, observations=obs, parameters=params.copy())
    for i, (start, S, B) in enumerate([(50, 16, 8), (0, 16, 8), (184, 16, 8), (80, 40, 10)]):
        T = obs.shape[0]
        end = start + S
        weights = np.linspace(1.0, 2.0, S)
        bd = dict(subsequence_start=start, subsequence_end=end, left_buffer_start=max(0, start - B),
                  right_buffer_end=min(T, end + B), weights=weights)
        g = sampler._single_noisy_grad_loglikelihood(buffer_dict=bd, kind="marginal")
        put("a/lgssm_%d" % i, obs=obs[bd["left_buffer_start"]:bd["right_buffer_end"]],
            t1=start - bd["left_buffer_start"], tL=end - bd["left_buffer_start"], weights=weights,
            theta=theta_of("lgssm", params), keys=np.array(sorted(g)),
            values=np.concatenate([np.ravel(g[k]) for k in sorted(g)]))
        print("a/lgssm_%d" % i, g, flush=True)


if __name__ == "__main__":
    data = kernel_level()
    helper_level(data)
    sampler_level(data)
    analytic_level(data)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, len(store), "arrays", os.path.getsize(OUT), "bytes")
