"""Drop-in for sgmcmc_ssm/particle_filters/buffered_smoother.py: same entry points and return dict,
with the whole time loop executed by the CUDA library (one C-ABI call per batch of work items).

  buffered_pf_wrapper(pf=, observations=, parameters=, N=, kernel=, additive_statistic_func=,
                      statistic_dim=, t1=, tL=, weights=, prior_mean=, prior_var=, **kwargs) -> dict
      keys x_t (N, n), log_weights (N,), statistics (N, p) [(p,) for pf='filter'],
      loglikelihood_estimate (float)                       (buffered_smoother.py:12-149, 156-199)
  average_statistic(out) -> (p,)                           (buffered_smoother.py:151-154)
  batched_pf(...)  -- the batched form used by the samplers (new; one launch for a whole minibatch)

Unknown kwargs are ignored, as in the reference (they flow through every layer there).  The
additive statistic must be one of the library's named statistics (the model score functions or the
sufficient statistics); arbitrary Python callables cannot run on the device.
"""
import numpy as np

from .. import engine
from .. import _native as nat
from . import statistics as S

_PF_NAMES = ("nemeth", "poyiadjis_N", "poyiadjis_N2", "paris", "filter")
_ENGINE_KW = ("dtype", "rng", "resample", "lambduh", "Ntilde", "accept_reject", "max_accept_reject",
              "manual_sample_threshold", "seed", "offset", "device", "item_id_base", "n2_mode", "num_steps_ahead", "per_horizon", "variates", "path")


_LAST_INJECTED = {}        # randoms drawn by the last rng='injected' call without explicit arrays (save_all replays prefixes)


def _theta(model, parameters):
    return S.MODEL_SPECS[model]["theta"](parameters)


def _draw_injected(N, T_list, pred=None):
    """Consume the GLOBAL numpy legacy stream exactly as the reference would for these items
    (SURVEY Appendix B): per item: N normals (sample_x0), then per step N uniforms (np.random.choice)
    and N normals (kernel.rv).  pred = (K, t1s, tLs): additionally the normal blocks of the predictive
    statistic, one per horizon k <= K with t + k < T, for steps in [t1, tL) (svm/helper.py:379,
    garch/helper.py:410 -> garch/kernels.py:66)."""
    B, max_T = len(T_list), max(T_list)
    z0 = np.zeros((B, N))
    u = np.zeros((B, max_T, N))
    z = np.zeros((B, max_T, N))
    zp = np.zeros((B, max_T, nat.PRED_SLOTS, N)) if pred is not None else None
    for b, T in enumerate(T_list):
        z0[b] = np.random.normal(size=N)
        for t in range(T):
            u[b, t] = np.random.random_sample(N)
            z[b, t] = np.random.normal(size=N)
            if pred is not None and pred[1][b] <= t < pred[2][b]:
                for k in range(pred[0] + 1):
                    if t + k >= T:
                        break
                    zp[b, t, k] = np.random.normal(size=N)
    out = dict(z0=z0, u=u, z=z)
    if zp is not None:
        out["zp"] = zp
    return out


def batched_pf(pf, model, kernel, items, N, stat_kind="score", want=(), sync=True, **kwargs):
    """Run one batch of work items; returns engine.PFResult."""
    if pf not in _PF_NAMES:
        raise ValueError("Unrecognized pf = {0}".format(pf))       # buffered_smoother.py:198
    kw = {k: kwargs[k] for k in _ENGINE_KW if k in kwargs and kwargs[k] is not None}
    rng = kw.get("rng", engine.config.rng)
    if rng == "injected" and "injected" not in kwargs:
        if pf == "paris":
            raise NotImplementedError("rng='injected' with pf='paris' needs a recorded stream "
                                      "(data-dependent number of draws); pass injected=dict(...)")
        pk = items.pack()
        pred = None
        if stat_kind == "pred" and model != "lgssm":      # the LGSSM predictive statistic is analytic (no draws)
            pred = (int(kw.get("num_steps_ahead", 5)), [int(v) for v in pk.t1], [int(v) for v in pk.tL])
        kw["injected"] = _LAST_INJECTED["value"] = _draw_injected(int(N), [int(T) for T in pk.T_buf], pred=pred)
        kw.setdefault("resample", "multinomial")
    elif "injected" in kwargs:
        kw["injected"] = kwargs["injected"]
    return engine.run_pf(model, kernel, pf, items, N, stat_kind=stat_kind, want=want, sync=sync,
                         while_running=kwargs.get("while_running"), **kw)


def buffered_pf_wrapper(pf, observations=None, parameters=None, N=1000, kernel=None,
                        additive_statistic_func=None, statistic_dim=None, t1=0, tL=None, weights=None,
                        prior_mean=0.0, prior_var=1.0, save_all=False, elementwise_statistic=False,
                        **kwargs):
    if pf not in _PF_NAMES:
        raise ValueError("Unrecognized pf = {0}".format(pf))
    if kernel is None or getattr(kernel, "model", None) is None:
        raise ValueError("kernel must be one of the library's Kernel objects")
    if kernel.kernel == "highdim":
        raise NotImplementedError("n > 1 latent states are outside the CUDA path")
    kernel.set_parameters(parameters=parameters)
    model = kernel.model
    stat_kind = S.stat_kind_of(additive_statistic_func)
    observations = np.asarray(observations, dtype=float)
    T = observations.shape[0]
    tL = T if tL is None else tL
    items = engine.PFItems().add(observations, _theta(model, parameters), t1=t1, tL=tL, weights=weights,
                                 prior_mean=float(np.ravel(prior_mean)[0]), prior_var=float(np.ravel(prior_var)[0]))
    if elementwise_statistic:
        return S.elementwise_run(pf, model, kernel.kernel, items, N, stat_kind, t1, tL, **kwargs)
    want = ("x", "lw") + (() if pf == "filter" else ("stats",)) + (("trace_x", "trace_lw") if save_all else ())
    res = batched_pf(pf, model, kernel.kernel, items, N, stat_kind=stat_kind, want=want, **kwargs)
    p = res.p
    out = dict(x_t=res.tensor("x")[0].double().cpu().numpy(),
               log_weights=res.tensor("lw")[0].double().cpu().numpy(),
               loglikelihood_estimate=float(res.loglik[0]))
    if pf == "filter":
        out["statistics"] = res.grad[0].copy()
    else:
        out["statistics"] = res.tensor("stats")[0][:, :p].double().cpu().numpy()
        out["_average_statistic"] = res.grad[0].copy()
    if save_all:
        # buffered_smoother.py:128-147: the particle system, statistics and running log-likelihood after every step
        out["all_x_t"] = res.tensor("trace_x")[0].double().cpu().numpy()
        out["all_log_weights"] = alw = res.tensor("trace_lw")[0].double().cpu().numpy()
        wts = None if weights is None else np.asarray(weights, dtype=float).reshape(-1)
        ll, all_ll = 0.0, [0.0]
        for t in range(T):
            if t1 <= t < tL:          # :124-126 (here max-shifted like the kernels)
                m = np.max(alw[t + 1])
                ll += (1.0 if wts is None else wts[t - t1]) * (m + np.log(np.mean(np.exp(alw[t + 1] - m))))
            all_ll.append(ll)
        out["all_loglikelihood_estimate"] = np.array(all_ll)
        # The statistics after step t are what a run over the first t observations returns: with the same random
        # numbers (Philox seed / call offset, or the recorded stream) the particle system of a prefix is identical.
        kw2 = dict(kwargs)
        if kw2.get("rng", engine.config.rng) == "injected":
            inj = kw2.get("injected") or _LAST_INJECTED.get("value")
        else:
            inj = None
            kw2.update(seed=int(res.seed), offset=int(res.offset))
        zero = np.zeros(p) if pf == "filter" else np.zeros((int(N), p))
        all_stats = [zero]
        for t in range(1, T + 1):
            sub = engine.PFItems().add(observations[:t], _theta(model, parameters), t1=min(t1, t), tL=min(tL, t),
                                       weights=weights, prior_mean=float(np.ravel(prior_mean)[0]),
                                       prior_var=float(np.ravel(prior_var)[0]))
            if inj is not None:
                kw2["injected"] = {k: (np.asarray(v)[:, :t] if k in ("u", "z", "zp") else
                                       (list(v)[:t] if k == "extra" else v)) for k, v in inj.items() if v is not None}
            r2 = batched_pf(pf, model, kernel.kernel, sub, N, stat_kind=stat_kind,
                            want=() if pf == "filter" else ("stats",), **kw2)
            all_stats.append(r2.grad[0].copy() if pf == "filter" else r2.tensor("stats")[0][:, :p].double().cpu().numpy())
        out["all_statistics"] = np.array(all_stats)
    return out


def pf_wrapper(observations, parameters, N, kernel, smoother, additive_statistic_func, statistic_dim,
               **kwargs):
    """buffered_smoother.py:12-149 signature; `smoother` is one of the named functions of
    particle_filters.pf (they carry the `pf` string)."""
    pf = getattr(smoother, "pf_name", None)
    if pf is None:
        raise ValueError("smoother must be one of particle_filters.pf.{nemeth,poyiadjis,paris}_smoother / pf_filter")
    if pf == "nemeth" and kwargs.get("lambduh") == 1.0:
        pf = "poyiadjis_N"
    return buffered_pf_wrapper(pf, observations=observations, parameters=parameters, N=N, kernel=kernel,
                               additive_statistic_func=additive_statistic_func, statistic_dim=statistic_dim,
                               **kwargs)


def average_statistic(out):
    """sum_i statistics[i, :] * softmax(log_weights)[i]   (buffered_smoother.py:151-154).
    The device already reduced it (fused into the last step); recompute only for foreign dicts."""
    if "_average_statistic" in out:
        return out["_average_statistic"]
    lw = np.asarray(out["log_weights"])
    w = np.exp(lw - np.max(lw))
    w /= np.sum(w)
    return np.sum(np.asarray(out["statistics"]).T * w, axis=1)
