"""Model helpers: the drop-in boundary `Helper.pf_gradient_estimate(...)` and its siblings.

Mirrors the PF entry points of sgmcmc_ssm/models/{svm,garch,lgssm}/helper.py (same names, argument
meaning, defaults, error behaviour, return keys); `*_batch` variants take many work items and issue
ONE C-ABI call.  Analytic message passing (LGSSM Kalman recursions) is outside the hot path.
"""
import numpy as np

from . import engine
from .particle_filters import statistics as S
from .particle_filters.buffered_smoother import batched_pf, buffered_pf_wrapper, average_statistic  # noqa: F401


class SGMCMCHelper(object):
    """Abstract helper (sgmcmc_sampler.py:1427-1964): only the particle-filter members are provided."""
    _model = None

    def _get_kernel(self, kernel):
        raise NotImplementedError()

    def _prior_moments(self, forward_message, parameters):
        """prior_var = inv(precision); prior_mean = solve(prior_var, mean_precision)  (sic)
        (svm/helper.py:100-104, lgssm/helper.py:1098-1101)."""
        if forward_message is None:
            forward_message = self.default_forward_message
        precision = np.atleast_2d(forward_message["precision"])
        if precision.size == 1:             # n = 1: inv / solve are one division each (bitwise what LAPACK returns)
            prior_var = 1.0 / float(precision[0, 0])
            return float(np.ravel(forward_message["mean_precision"])[0]) / prior_var, prior_var
        prior_var = np.linalg.inv(precision)
        prior_mean = np.linalg.solve(prior_var, np.atleast_1d(forward_message["mean_precision"]))
        return float(prior_mean[0]), float(prior_var[0, 0])

    # ---- batched form ---------------------------------------------------------------------------
    def make_items(self, windows, parameters, forward_message=None):
        """windows: iterable of dicts(observations, subsequence_start, subsequence_end, weights)."""
        prior_mean, prior_var = self._prior_moments(forward_message, parameters)
        theta = S.MODEL_SPECS[self._model]["theta"](parameters)
        items = engine.PFItems()
        for w in windows:
            items.add(w["observations"], theta, t1=w.get("subsequence_start", 0), tL=w.get("subsequence_end"),
                      weights=w.get("weights"), prior_mean=prior_mean, prior_var=prior_var)
        return items

    def pf_gradient_estimate_batch(self, windows, parameters, pf="poyiadjis_N", N=1000, kernel=None,
                                   forward_message=None, sync=True, **kwargs):
        """List of gradient dicts (one per window) + the PFResult; one device launch sequence."""
        K = self._get_kernel(kernel)
        K.set_parameters(parameters)
        if N is None:
            raise TypeError("N (number of particles) must be given for kind='pf'")
        items = self.make_items(windows, parameters, forward_message)
        res = batched_pf(pf, K.model, K.kernel, items, N, stat_kind="score", sync=sync, **kwargs)
        if not sync:
            return res
        keys = S.MODEL_SPECS[self._model]["grad_keys"]
        return [{k: g[i] for i, k in enumerate(keys)} for g in res.grad], res

    def pf_gradient_sum_packed(self, packed, parameters, pf="poyiadjis_N", N=1000, kernel=None, **kwargs):
        """Sum over the items of an engine.PackedItems batch of their gradient estimates (dict keyed like
        pf_gradient_estimate) -- the vectorised path of the samplers: no per-item Python work."""
        K = self._get_kernel(kernel)
        K.set_parameters(parameters)
        if N is None:
            raise TypeError("N (number of particles) must be given for kind='pf'")
        keys = S.MODEL_SPECS[self._model]["grad_keys"]
        rng = kwargs.get("rng", engine.config.rng)
        if rng == "injected":          # parity mode: the host draws the recorded stream item by item
            res = batched_pf(pf, K.model, K.kernel, packed, N, stat_kind="score", **kwargs)
            total = res.grad.sum(axis=0)
            return {k: float(total[i]) for i, k in enumerate(keys)}, res
        from .particle_filters.buffered_smoother import _ENGINE_KW
        kw = {k: kwargs[k] for k in _ENGINE_KW if k in kwargs and kwargs[k] is not None}
        total, info = engine.run_pf_sum(K.model, K.kernel, pf, packed, N, allreduce=kwargs.get("allreduce", False),
                                        while_running=kwargs.get("while_running"), stat_kind="score", **kw)
        return {k: float(total[i]) for i, k in enumerate(keys)}, info

    def packed_items(self, parameters, forward_message=None, **arrays):
        prior_mean, prior_var = self._prior_moments(forward_message, parameters)
        theta = S.MODEL_SPECS[self._model]["theta"](parameters)
        return engine.PackedItems(theta=theta, prior_mean=prior_mean, prior_var=prior_var, **arrays)

    def pf_loglikelihood_estimate_batch(self, windows, parameters, pf="poyiadjis_N", N=1000, kernel=None,
                                        forward_message=None, **kwargs):
        K = self._get_kernel(kernel)
        K.set_parameters(parameters)
        items = self.make_items(windows, parameters, forward_message)
        res = batched_pf(pf, K.model, K.kernel, items, N, stat_kind="none", **kwargs)
        return res.loglik.copy()

    # ---- reference-shaped single-item entry points ------------------------------------------------
    def pf_gradient_estimate(self, observations, parameters, subsequence_start=0, subsequence_end=None,
                             weights=None, pf="poyiadjis_N", N=1000, kernel=None, forward_message=None,
                             **kwargs):
        """Particle-filter score estimate (svm/helper.py:67-128, garch/helper.py:59-117,
        lgssm/helper.py:1089-1143).  Returns dict of gradients keyed like the reference."""
        grads, _ = self.pf_gradient_estimate_batch(
            [dict(observations=observations, subsequence_start=subsequence_start,
                  subsequence_end=subsequence_end, weights=weights)],
            parameters, pf=pf, N=N, kernel=kernel, forward_message=forward_message, **kwargs)
        return grads[0]

    def pf_loglikelihood_estimate(self, observations, parameters, subsequence_start=0, subsequence_end=None,
                                  weights=None, pf="poyiadjis_N", N=1000, kernel=None, forward_message=None,
                                  **kwargs):
        """svm/helper.py:130-185."""
        return float(self.pf_loglikelihood_estimate_batch(
            [dict(observations=observations, subsequence_start=subsequence_start,
                  subsequence_end=subsequence_end, weights=weights)],
            parameters, pf=pf, N=N, kernel=kernel, forward_message=forward_message, **kwargs)[0])

    def pf_latent_var_distr(self, observations, parameters, lag=None, subsequence_start=0,
                            subsequence_end=None, weights=None, pf="poyiadjis_N", N=1000, kernel=None,
                            forward_message=None, squared=False, **kwargs):
        """Smoothed latent marginals (svm/helper.py:249-294, garch/helper.py pf_latent_var_distr)."""
        if lag == 0 and pf != "filter":
            raise ValueError("pf must be filter for lag = 0")
        elif lag is None and pf == "filter":
            raise ValueError("pf must not be filter for smoothing")
        elif lag is not None and lag != 0:
            raise NotImplementedError("lag can only be None or 0")
        K = self._get_kernel(kernel)
        prior_mean, prior_var = self._prior_moments(forward_message, parameters)
        stat = S.garch_sufficient_statistics if self._model == "garch" else S.gaussian_sufficient_statistics
        out = buffered_pf_wrapper(pf=pf, observations=observations, parameters=parameters, N=N, kernel=K,
                                  additive_statistic_func=stat, statistic_dim=3, t1=subsequence_start,
                                  tL=subsequence_end, weights=weights, prior_mean=prior_mean,
                                  prior_var=prior_var, elementwise_statistic=True, **kwargs)
        # lag = 0: the filter's statistic is already the weighted average (pf.py:77-80)
        avg = np.reshape(out["statistics"] if pf == "filter" else average_statistic(out), (-1, 3))
        if self._model == "garch" and squared:
            x_mean, x_cov = avg[:, 1], avg[:, 2] - avg[:, 1] ** 2
        else:
            x_mean, x_cov = avg[:, 0], avg[:, 1] - avg[:, 0] ** 2
        return np.reshape(x_mean, (x_mean.shape[0], 1)), np.reshape(x_cov, (x_cov.shape[0], 1, 1))

    def pf_predictive_loglikelihood_estimate(self, observations, parameters, num_steps_ahead=5,
                                             subsequence_start=0, subsequence_end=None, pf="filter",
                                             N=1000, kernel=None, forward_message=None, **kwargs):
        """k-step-ahead predictive log-likelihood for k = 0..num_steps_ahead (svm/helper.py:187-247,
        lgssm/helper.py:1048-1087, garch/helper.py pf_predictive_loglikelihood_estimate): particle FILTER with
        the `logsumexp` statistic; element 0 is the log-likelihood estimate.  per_horizon=False (default)
        reproduces the reference's log-sum over ALL horizons (pf.py:73-76); True gives one log-sum per horizon."""
        return self.pf_predictive_loglikelihood_estimate_batch(
            [dict(observations=observations, subsequence_start=subsequence_start, subsequence_end=subsequence_end)],
            parameters, num_steps_ahead=num_steps_ahead, pf=pf, N=N, kernel=kernel, forward_message=forward_message,
            **kwargs)[0]

    def pf_predictive_loglikelihood_estimate_batch(self, windows, parameters, num_steps_ahead=5, pf="filter", N=1000,
                                                   kernel=None, forward_message=None, **kwargs):
        if pf != "filter":
            raise ValueError("Only can use pf = 'filter' since we are filtering")
        K = self._get_kernel(kernel)
        K.set_parameters(parameters)
        kwargs.pop("weights", None)                     # the reference passes no weights here
        items = self.make_items([dict(w, weights=None) for w in windows], parameters, forward_message)
        res = batched_pf(pf, K.model, K.kernel, items, N, stat_kind="pred", num_steps_ahead=num_steps_ahead, **kwargs)
        out = res.grad[:, :int(num_steps_ahead) + 1].copy()
        out[:, 0] = res.loglik
        return out

    def _forward_messages(self, *args, **kwargs):
        raise NotImplementedError("analytic message passing is outside the particle-filter hot path")

    _backward_messages = _forward_messages
