"""Named additive statistics and per-model glue (theta packing, gradient keys)."""
import numpy as np
import torch


def _f(a):
    return float(np.ravel(a)[0])


def _gauss_dims_ok(p):
    if getattr(p, "n", 1) != 1 or getattr(p, "m", 1) != 1:
        raise NotImplementedError("the CUDA particle-filter path covers n = m = 1 (reference scalar branches)")


def svm_theta(p):
    _gauss_dims_ok(p)
    # slots 10, 11: Q, R as the Parameters object computes them (predictive statistic, svm/helper.py:374)
    return [_f(p.A), _f(p.LQinv), _f(p.Qinv), _f(p.LRinv), _f(p.Rinv), 0.0, 0.0, 0.0, 0.0, 0.0, _f(p.Q), _f(p.R)]


def lgssm_theta(p):
    _gauss_dims_ok(p)
    return [_f(p.A), _f(p.LQinv), _f(p.Qinv), _f(p.C), _f(p.LRinv), _f(p.Rinv), 0.0, 0.0, 0.0, 0.0, _f(p.Q), _f(p.R)]


def garch_theta(p):
    return [_f(p.alpha), _f(p.beta), _f(p.gamma), _f(p.mu), _f(p.phi), _f(p.lambduh), _f(p.LRinv), _f(p.Rinv), _f(p.R)]


MODEL_SPECS = {
    "svm": dict(theta=svm_theta, grad_keys=("LRinv_vec", "LQinv_vec", "A")),                       # svm/helper.py:122-126
    "lgssm": dict(theta=lgssm_theta, grad_keys=("LRinv_vec", "LQinv_vec", "C", "A")),              # lgssm/helper.py:1137-1142
    "garch": dict(theta=garch_theta, grad_keys=("LRinv_vec", "log_mu", "logit_phi", "logit_lambduh")),  # garch/helper.py:110-115
}


def _marker(kind, name, doc):
    def stat(*args, **kwargs):
        raise NotImplementedError("%s is evaluated inside the CUDA kernels" % name)
    stat.stat_kind = kind
    stat.__name__ = name
    stat.__doc__ = doc
    return stat


svm_complete_data_loglike_gradient = _marker("score", "svm_complete_data_loglike_gradient", "svm/helper.py:297-350")
lgssm_complete_data_loglike_gradient = _marker("score", "lgssm_complete_data_loglike_gradient", "lgssm/helper.py:1217-1279")
garch_complete_data_loglike_gradient = _marker("score", "garch_complete_data_loglike_gradient", "garch/helper.py:335-372")
gaussian_sufficient_statistics = _marker("suff", "gaussian_sufficient_statistics", "lgssm/helper.py:1338-1363")
garch_sufficient_statistics = _marker("suff", "garch_sufficient_statistics", "garch/helper.py:414-434")


def stat_kind_of(func):
    if func is None:
        return "none"
    kind = getattr(func, "stat_kind", None)
    if kind is None:
        raise NotImplementedError("additive_statistic_func must be one of the library's named statistics "
                                  "(arbitrary Python callables cannot run on the device)")
    return kind


def elementwise_run(pf, model, kernel, items, N, stat_kind, t1, tL, **kwargs):
    """elementwise_statistic=True (buffered_smoother.py:106-112, 201-210): statistic h_s of every step
    s in [t1, tL) kept separate.  For the genealogy-tracking smoothers (poyiadjis_N / nemeth with
    lambduh = 1) the wide statistic of particle i is h_s evaluated along i's ancestral line, so it is
    rebuilt from the ancestor and particle traces (T gathers) instead of widening the records."""
    from .buffered_smoother import batched_pf
    if pf not in ("poyiadjis_N", "filter") or stat_kind != "suff":
        raise NotImplementedError("elementwise statistics are implemented for pf='poyiadjis_N' (smoothing) and "
                                  "pf='filter' (lag = 0) with the sufficient statistics (what pf_latent_var_distr uses)")
    if pf == "filter":
        # pf.py:40-82 with the wide statistic: block s of the (3 L,) vector is sum_i h_s(x_anc_i, x'_i) softmax(lw')_i of
        # step s, from the traced particle system
        res = batched_pf(pf, model, kernel, items, N, stat_kind="none", want=("lw", "anc", "trace_x", "trace_lw", "x"), **kwargs)
        anc = res.tensor("anc")[0].long()
        tx = res.tensor("trace_x")[0].double()
        tlw = res.tensor("trace_lw")[0].double()
        L = tL - t1
        stats = torch.zeros(3 * L, dtype=torch.float64, device=anc.device)
        wts = items.weights[0]
        for t in range(t1, min(tL, anc.shape[0])):
            xn, xa = tx[t + 1][:, 0], tx[t][anc[t]][:, 0]
            w = torch.softmax(tlw[t + 1], dim=0)
            s = t - t1
            h = torch.stack([xn, xn * xn, (xn ** 4) if model == "garch" else xa * xn])
            stats[3 * s:3 * s + 3] = (h * w).sum(dim=1) * (1.0 if wts is None else float(wts[s]))
        return dict(x_t=res.tensor("x")[0].double().cpu().numpy(), log_weights=res.tensor("lw")[0].double().cpu().numpy(),
                    statistics=stats.cpu().numpy(), loglikelihood_estimate=float(res.loglik[0]))
    res = batched_pf(pf, model, kernel, items, N, stat_kind="none", want=("lw", "anc", "trace_x", "x"), **kwargs)
    anc = res.tensor("anc")[0].long()            # (T, N)
    tx = res.tensor("trace_x")[0].double()       # (T + 1, N, n)
    lw = res.tensor("lw")[0].double()
    T = anc.shape[0]
    L = tL - t1
    stats = torch.zeros((N, 3 * L), dtype=torch.float64, device=anc.device)
    wts = items.weights[0]
    idx = torch.arange(N, device=anc.device)     # lineage index at time t + 1
    for t in range(T - 1, -1, -1):
        parent = anc[t][idx]
        if t1 <= t < tL:
            xn = tx[t + 1][idx][:, 0]
            xa = tx[t][parent][:, 0]
            s = t - t1
            stats[:, 3 * s] = xn
            stats[:, 3 * s + 1] = xn * xn
            stats[:, 3 * s + 2] = (xn ** 4) if model == "garch" else xa * xn
            if wts is not None:                  # additive_scale multiplies the statistic (pf.py:173)
                stats[:, 3 * s:3 * s + 3] *= float(wts[s])
        idx = parent
    out = dict(x_t=res.tensor("x")[0].double().cpu().numpy(), log_weights=lw.cpu().numpy(),
               statistics=stats.cpu().numpy(), loglikelihood_estimate=float(res.loglik[0]))
    return out
