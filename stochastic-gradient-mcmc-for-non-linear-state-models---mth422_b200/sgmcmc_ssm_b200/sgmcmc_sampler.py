"""SG-MCMC sampler: the caller of the particle-filter hot path.

API mirror of sgmcmc_ssm/sgmcmc_sampler.py (SGMCMCSampler :12-1157, SeqSGMCMCSampler :1159-1283,
random_subsequence_and_weights :1969-2017) for the `kind='pf'` path: same method names, arguments,
defaults, kwargs pass-through and error behaviour.  What changes underneath: all `minibatch_size`
subsequences of a gradient (and all sequences of a Seq sampler) are packed into ONE batched C-ABI
call instead of a Python loop over `pf_gradient_estimate`.

`kind='marginal'` / `'complete'` (analytic message passing, Gibbs) are outside the hot path and raise
NotImplementedError.
"""
import logging
import time

import numpy as np

logger = logging.getLogger(name=__name__)
NOISE_NUGGET = 1e-9


def random_subsequence_and_weights(S, T, partition_style=None):
    """Random subsequence [start, end) of length S and its per-step importance weights
    (sgmcmc_sampler.py:1969-2017; same numpy draws)."""
    if partition_style is None:
        partition_style = "uniform"
    if partition_style == "strict":
        if T % S != 0:
            raise ValueError("S {0} does not evenly divide T {1}".format(S, T))
        start = np.random.choice(np.arange(0, T // S)) * S
        end = start + S
        weights = np.ones(S, dtype=float) * T / S
    elif partition_style == "uniform":
        start = np.random.randint(0, T - S + 1)
        end = start + S
        t = np.arange(start, end)
        if end <= 2 * S:
            num = np.min(np.array([t + 1, np.ones_like(t) * min(S, T - S + 1)]), axis=0)
        elif start >= T - 2 * S - 1:
            num = np.min(np.array([T - t, np.ones_like(t) * min(S, T - S + 1)]), axis=0)
        else:
            num = np.ones(S) * S
        weights = np.ones(S, dtype=float) * (T - S + 1) / num
    elif partition_style == "naive":
        start = np.random.randint(0, T - S + 1)
        end = start + S
        weights = np.ones(S, dtype=float) * T / S
    else:
        raise ValueError("Unrecognized partition_style = '{0}'".format(partition_style))
    return int(start), int(end), weights


def random_subsequences_packed(observations, S, M, buffer_length, partition_style=None, lo=0, hi=None):
    """Vectorised form of M calls of random_subsequence_and_weights + the buffer / window slicing of
    sgmcmc_sampler.py:259-288, 364-374 for the 'uniform' and 'naive' partition styles.  Consumes the numpy
    stream exactly like the M sequential `np.random.randint(0, T - S + 1)` calls of the reference (legacy
    RandomState draws bounded integers element by element: checked in tests/test_host_logic.py).
    Returns the keyword arrays of engine.PackedItems (without theta / prior) for the items [lo, hi) of the minibatch."""
    T = observations.shape[0]
    if buffer_length == -1:
        buffer_length = T
    start = np.random.randint(0, T - S + 1, size=M).astype(np.int64)     # all M draws: every rank keeps the same stream
    start = start[lo:hi]                                                  # ... and packs only its own shard
    M = start.shape[0]
    end = start + S
    if partition_style in (None, "uniform"):
        t = start[:, None] + np.arange(S)[None, :]
        cap = min(S, T - S + 1)
        num = np.full((M, S), float(S))
        head = end <= 2 * S
        tail = (~head) & (start >= T - 2 * S - 1)
        num[head] = np.minimum(t[head] + 1, cap)
        num[tail] = np.minimum(T - t[tail], cap)
        weights = (T - S + 1) / num
    elif partition_style == "naive":
        weights = np.full((M, S), float(T) / S)
    else:
        raise ValueError("Unrecognized partition_style = '{0}'".format(partition_style))
    left = np.maximum(0, start - buffer_length)
    right = np.minimum(T, end + buffer_length)
    lens = right - left
    off = np.concatenate([[0], np.cumsum(lens)[:-1]])
    idx = np.arange(int(lens.sum())) - np.repeat(off, lens) + np.repeat(left, lens)
    obs = np.asarray(observations, dtype=np.float64).reshape(T, -1)
    return dict(obs_flat=obs[idx, 0], T_buf=lens.astype(np.int32), t1=(start - left).astype(np.int32),
                tL=(end - left).astype(np.int32), wts_flat=weights.reshape(-1),
                wts_off=np.arange(M, dtype=np.int64) * S)


def _window(observations, bd):
    """Buffered window + relative subsequence of one buffer_dict (sgmcmc_sampler.py:364-374)."""
    return dict(observations=observations[bd["left_buffer_start"]:bd["right_buffer_end"]],
                subsequence_start=bd["subsequence_start"] - bd["left_buffer_start"],
                subsequence_end=bd["subsequence_end"] - bd["left_buffer_start"],
                weights=bd["weights"])


def _merge_injected(drawn):
    """Concatenate per-item recorded randoms along the batch axis (time axis padded)."""
    max_T = max(d["u"].shape[1] for d in drawn)

    def pad(a):
        return np.pad(a, ((0, 0), (0, max_T - a.shape[1]), (0, 0)))
    return dict(z0=np.concatenate([d["z0"] for d in drawn]), u=np.concatenate([pad(d["u"]) for d in drawn]),
                z=np.concatenate([pad(d["z"]) for d in drawn]))


_NOT_PF = "kind='{0}' (analytic message passing) is outside the particle-filter hot path; use kind='pf'"


class SGMCMCSampler(object):
    """Base class for SG-MCMC on time series (subclasses set prior / parameters / message_helper)."""

    def __init__(self, **kwargs):
        raise NotImplementedError()

    def prior_init(self):
        self.parameters = self.prior.sample_prior()
        return self.parameters

    # ---- log-likelihood -------------------------------------------------------------------------
    def noisy_loglikelihood(self, kind="marginal", subsequence_length=-1, minibatch_size=1,
                            buffer_length=10, num_samples=None, observations=None, **kwargs):
        """Subsequence approximation of the log-likelihood (sgmcmc_sampler.py:130-243, pf branch)."""
        observations = self._get_observations(observations)
        T = observations.shape[0]
        if kind != "pf":
            if kind in ("marginal", "complete"):
                raise NotImplementedError(_NOT_PF.format(kind))
            raise ValueError("Unrecognized kind = {0}".format(kind))
        if kwargs.get("N", None) is None:
            kwargs["N"] = num_samples
        from . import engine
        from .particle_filters.buffered_smoother import _draw_injected
        kwargs.pop("parameters", None)
        replay = kwargs.get("rng", engine.config.rng) == "injected" and "injected" not in kwargs
        windows, drawn = [], []
        for _ in range(minibatch_size):
            # the reference draws a subsequence and runs its filter before drawing the next one (:214-237)
            bd = self._random_subsequence_and_buffers(buffer_length=buffer_length,
                                                      subsequence_length=subsequence_length, T=T)
            windows.append(_window(observations, bd))
            if replay:
                drawn.append(_draw_injected(int(kwargs["N"]), [windows[-1]["observations"].shape[0]]))
        if replay:
            kwargs["injected"] = _merge_injected(drawn)
            kwargs.setdefault("resample", "multinomial")
        ll = self.message_helper.pf_loglikelihood_estimate_batch(windows, self.parameters, **kwargs)
        noisy_loglikelihood = float(np.sum(ll)) * 1.0 / minibatch_size
        if np.isnan(noisy_loglikelihood):
            raise ValueError("NaNs in loglikelihood")
        return noisy_loglikelihood

    def noisy_logjoint(self, return_loglike=False, **kwargs):
        loglikelihood = self.noisy_loglikelihood(**kwargs)
        logprior = self.prior.logprior(self.parameters)
        if return_loglike:
            return dict(logjoint=loglikelihood + logprior, loglikelihood=loglikelihood)
        return loglikelihood + logprior

    def predictive_loglikelihood(self, kind="marginal", num_steps_ahead=10, subsequence_length=-1, minibatch_size=1,
                                 buffer_length=10, num_samples=1000, parameters=None, observations=None, **kwargs):
        """Subsequence estimate of the k-step-ahead predictive log-likelihood, k = 0..num_steps_ahead
        (sgmcmc_sampler.py:61-128, pf branch :94-126): every minibatch item's estimate is rescaled by
        (T - k) / (S - k); all items run in one batched launch."""
        observations = self._get_observations(observations)
        T = observations.shape[0]
        if kind != "pf":
            if kind in ("marginal", "complete"):
                raise NotImplementedError(_NOT_PF.format(kind))
            raise ValueError("Unrecognized kind = {0}".format(kind))
        if kwargs.get("N", None) is None:
            kwargs["N"] = num_samples
        from . import engine
        from .particle_filters.buffered_smoother import _draw_injected
        replay = kwargs.get("rng", engine.config.rng) == "injected" and "injected" not in kwargs
        windows, scales, drawn = [], [], []
        for _ in range(minibatch_size):
            bd = self._random_subsequence_and_buffers(buffer_length=buffer_length,
                                                      subsequence_length=subsequence_length, T=T)
            windows.append(_window(observations, bd))
            S = bd["subsequence_end"] - bd["subsequence_start"]
            with np.errstate(divide="ignore"):
                scales.append((T - np.arange(num_steps_ahead + 1)) / (S - np.arange(num_steps_ahead + 1.0)))
            if replay:                                  # the reference filters each window before drawing the next
                w = windows[-1]
                pred = None if self.message_helper._model == "lgssm" else (
                    num_steps_ahead, [w["subsequence_start"]], [w["subsequence_end"]])
                drawn.append(_draw_injected(int(kwargs["N"]), [w["observations"].shape[0]], pred=pred))
        if replay:
            merged = _merge_injected(drawn)
            if "zp" in drawn[0]:
                max_T = merged["u"].shape[1]
                merged["zp"] = np.concatenate([np.pad(d["zp"], ((0, 0), (0, max_T - d["zp"].shape[1]), (0, 0), (0, 0)))
                                               for d in drawn])
            kwargs["injected"] = merged
            kwargs.setdefault("resample", "multinomial")
        est = self.message_helper.pf_predictive_loglikelihood_estimate_batch(
            windows, self.parameters, num_steps_ahead=num_steps_ahead, **kwargs)
        pred_loglikelihood = np.zeros(num_steps_ahead + 1)
        for add, sc in zip(est, scales):
            pred_loglikelihood += add * sc
        return pred_loglikelihood * 1.0 / minibatch_size

    # ---- gradients ------------------------------------------------------------------------------
    def _random_subsequence_and_buffers(self, buffer_length, subsequence_length, T=None):
        """sgmcmc_sampler.py:259-288."""
        if T is None:
            T = self._get_T()
        if buffer_length == -1:
            buffer_length = T
        if (subsequence_length == -1) or (T - subsequence_length <= 0):
            subsequence_start, subsequence_end, weights = 0, T, None
        else:
            subsequence_start, subsequence_end, weights = random_subsequence_and_weights(
                S=subsequence_length, T=T, partition_style=self.options.get("partition_style"))
        return dict(subsequence_start=subsequence_start, subsequence_end=subsequence_end,
                    left_buffer_start=max(0, subsequence_start - buffer_length),
                    right_buffer_end=min(T, subsequence_end + buffer_length), weights=weights)

    def _single_noisy_grad_loglikelihood(self, buffer_dict, kind="marginal", num_samples=None,
                                         observations=None, parameters=None, **kwargs):
        """One subsequence (sgmcmc_sampler.py:290-388).  Note the reference's pf branch always uses
        self.parameters (:379); reproduced."""
        observations = self._get_observations(observations, check_shape=False)
        if kind != "pf":
            if kind in ("marginal", "complete"):
                raise NotImplementedError(_NOT_PF.format(kind))
            raise ValueError("Unrecognized kind = {0}".format(kind))
        if kwargs.get("N", None) is None:
            kwargs["N"] = num_samples
        w = _window(observations, buffer_dict)
        return self.message_helper.pf_gradient_estimate(
            observations=w["observations"], parameters=self.parameters,
            subsequence_start=w["subsequence_start"], subsequence_end=w["subsequence_end"],
            weights=w["weights"], **kwargs)

    def _pf_windows(self, subsequence_length=-1, minibatch_size=1, buffer_length=0, observations=None,
                    buffer_dicts=None):
        observations = self._get_observations(observations, check_shape=False)
        T = observations.shape[0]
        if buffer_dicts is None:
            buffer_dicts = [self._random_subsequence_and_buffers(buffer_length=buffer_length,
                                                                 subsequence_length=subsequence_length, T=T)
                            for _ in range(minibatch_size)]
        elif len(buffer_dicts) != minibatch_size:
            raise ValueError("len(buffer_dicts != minibatch_size")
        return [_window(observations, bd) for bd in buffer_dicts]

    def _pf_plan(self, subsequence_length=-1, minibatch_size=1, buffer_length=0, observations=None, buffer_dicts=None,
                 **unused):
        """The work items of one gradient evaluation and how to combine their gradients: (windows, finish) with
        finish(list of per-window gradient dicts) -> noisy_grad dict (sgmcmc_sampler.py:390-425).  Lets
        ensemble.ChainEnsemble pack the items of many chains into one launch."""
        windows = self._pf_windows(subsequence_length, minibatch_size, buffer_length, observations, buffer_dicts)

        def finish(grads):
            noisy_grad = {var: np.zeros_like(value) for var, value in self.parameters.as_dict().items()}
            for g in grads:
                for var in noisy_grad:
                    noisy_grad[var] += g[var] * 1.0 / minibatch_size
            return self._check_noisy_grad(noisy_grad)
        return windows, finish

    def _noisy_grad_loglikelihood(self, subsequence_length=-1, minibatch_size=1, buffer_length=0,
                                  observations=None, buffer_dicts=None, kind="marginal", num_samples=None,
                                  parameters=None, **kwargs):
        """Minibatch mean of subsequence gradients (sgmcmc_sampler.py:390-425) -- all `minibatch_size`
        particle filters run in one batched launch."""
        if kind != "pf":
            if kind in ("marginal", "complete"):
                raise NotImplementedError(_NOT_PF.format(kind))
            raise ValueError("Unrecognized kind = {0}".format(kind))
        if kwargs.get("N", None) is None:
            kwargs["N"] = num_samples
        noisy_grad = {var: np.zeros_like(value) for var, value in self.parameters.as_dict().items()}
        obs_all = self._get_observations(observations, check_shape=False)
        T_all = obs_all.shape[0]
        style = self.options.get("partition_style")
        if (buffer_dicts is None and minibatch_size > 1 and subsequence_length != -1 and T_all - subsequence_length > 0
                and style in (None, "uniform", "naive") and np.ndim(obs_all) == 2 and obs_all.shape[1] == 1):
            # vectorised minibatch: same numpy draws as the loop below, no per-item Python work
            from . import parallel
            distributed = kwargs.pop("distributed", False)
            lo, hi = parallel.shard_bounds(minibatch_size) if distributed else (0, minibatch_size)
            fm = kwargs.pop("forward_message", None)
            arrays = random_subsequences_packed(obs_all, subsequence_length, minibatch_size, buffer_length, style, lo, hi)
            keys = list(noisy_grad)
            from . import engine
            if kwargs.get("rng", engine.config.rng) == "injected":
                local = np.zeros(len(keys))
                if hi > lo:
                    packed = self.message_helper.packed_items(self.parameters, forward_message=fm, **arrays)
                    sums, _ = self.message_helper.pf_gradient_sum_packed(packed, self.parameters, item_id_base=lo, **kwargs)
                    local = np.array([sums[k] for k in keys])
                total = parallel.allreduce_sum(local) if distributed else local
            else:
                # device randoms: the item gradients are summed on the device, all-reduced in place over the ranks
                # (one NCCL call on the same buffer) and read back with a single small D2H copy
                packed = self.message_helper.packed_items(self.parameters, forward_message=fm, **arrays) if hi > lo else None
                sums, info = self.message_helper.pf_gradient_sum_packed(packed, self.parameters, item_id_base=lo,
                                                                        allreduce=distributed, **kwargs)
                self.last_pf_info = info
                total = np.array([sums[k] for k in keys])
                if distributed and not (isinstance(info, dict) and info.get("allreduced")):
                    total = parallel.allreduce_sum(total)
            for k, v in zip(keys, total):
                noisy_grad[k] += v / minibatch_size
            return self._check_noisy_grad(noisy_grad)
        windows = self._pf_windows(subsequence_length, minibatch_size, buffer_length, observations, buffer_dicts)
        if kwargs.pop("distributed", False):
            # every rank holds the same sampler state and numpy stream, hence the same windows; rank r
            # filters its contiguous shard and ONE all-reduce sums the per-rank gradient sums
            from . import parallel
            lo, hi = parallel.shard_bounds(len(windows))
            keys = list(noisy_grad)
            local = np.zeros(len(keys))
            if hi > lo:
                grads, _ = self.message_helper.pf_gradient_estimate_batch(
                    windows[lo:hi], self.parameters, item_id_base=lo, **kwargs)
                local = np.array([sum(float(np.ravel(g[k])[0]) for g in grads) for k in keys])
            else:
                from . import engine
                engine.skip_call()
            total = parallel.allreduce_sum(local)
            for k, v in zip(keys, total):
                noisy_grad[k] += v / minibatch_size
        else:
            grads, _ = self.message_helper.pf_gradient_estimate_batch(windows, self.parameters, **kwargs)
            for g in grads:
                for var in noisy_grad:
                    noisy_grad[var] += g[var] * 1.0 / minibatch_size
        return self._check_noisy_grad(noisy_grad)

    @staticmethod
    def _check_noisy_grad(noisy_grad):
        for var in noisy_grad:
            if np.any(np.isnan(noisy_grad[var])):
                raise ValueError("NaNs in gradient of {0}".format(var))
            g = noisy_grad[var]
            if (abs(g.flat[0]) if np.size(g) == 1 else np.linalg.norm(g)) > 1e16:
                logger.warning("Norm of noisy_grad_loglike[{1} > 1e16: {0}".format(noisy_grad[var], var))
        return noisy_grad

    def noisy_gradient(self, preconditioner=None, is_scaled=True, **kwargs):
        """grad log-likelihood estimate + grad log-prior, scaled by 1/T (sgmcmc_sampler.py:427-464)."""
        noisy_grad_loglike = kwargs.pop("noisy_grad_loglike", None)       # precomputed by ensemble.ChainEnsemble
        if noisy_grad_loglike is None and kwargs.get("kind") == "pf" and not kwargs.get("distributed", False):
            # the prior gradient (host numpy, no random numbers) is evaluated while the particle filters run
            box = []
            noisy_grad_loglike = self._noisy_grad_loglikelihood(
                while_running=lambda: box or box.append(
                    self.prior.grad_logprior(parameters=kwargs.get("parameters", self.parameters))),
                **kwargs)
            noisy_grad_prior = box[0] if box else self.prior.grad_logprior(parameters=kwargs.get("parameters", self.parameters))
        else:
            if noisy_grad_loglike is None:
                noisy_grad_loglike = self._noisy_grad_loglikelihood(**kwargs)
            noisy_grad_prior = self.prior.grad_logprior(parameters=kwargs.get("parameters", self.parameters))
        noisy_gradient = {var: noisy_grad_prior[var] + noisy_grad_loglike[var] for var in noisy_grad_prior}
        if preconditioner is None:
            if is_scaled:
                for var in noisy_gradient:
                    noisy_gradient[var] /= self._get_T(**kwargs)
        else:
            scale = 1.0 / self._get_T(**kwargs) if is_scaled else 1.0
            noisy_gradient = preconditioner.precondition(
                noisy_gradient, parameters=kwargs.get("parameters", self.parameters), scale=scale)
        return noisy_gradient

    # ---- steps ----------------------------------------------------------------------------------
    def step_sgd(self, epsilon, **kwargs):
        delta = self.noisy_gradient(**kwargs)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += epsilon * delta[var]
        return self.parameters

    def step_precondition_sgd(self, epsilon, preconditioner, **kwargs):
        delta = self.noisy_gradient(preconditioner=preconditioner, **kwargs)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += epsilon * delta[var]
        return self.parameters

    def step_adagrad(self, epsilon, **kwargs):
        if not hasattr(self, "_adagrad_moments"):
            self._adagrad_moments = dict(t=0, G=0.0)
        g = self.parameters.from_dict_to_vector(self.noisy_gradient(**kwargs))
        G = self._adagrad_moments["G"] + g ** 2
        delta = self.parameters.from_vector_to_dict(g / np.sqrt(G + NOISE_NUGGET), **self.parameters.dim)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += epsilon * delta[var]
        self._adagrad_moments = dict(t=self._adagrad_moments["t"] + 1, G=G)
        return self.parameters

    def _get_sgmcmc_noise(self, is_scaled=True, preconditioner=None, **kwargs):
        """N(0, 1/T) per parameter, var_dict order (sgmcmc_sampler.py:529-547)."""
        scale = 1.0 / self._get_T(**kwargs) if is_scaled else 1.0
        if preconditioner is not None:
            return preconditioner.precondition_noise(parameters=self.parameters, scale=scale)
        return {var: np.random.normal(loc=0, scale=np.sqrt(scale), size=value.shape)
                for var, value in self.parameters.as_dict().items()}

    def sample_sgld(self, epsilon, **kwargs):
        """theta += eps * grad + sqrt(2 eps) * N(0, 1/T)   (sgmcmc_sampler.py:549-567)."""
        if "preconditioner" in kwargs:
            raise ValueError("Use SGRLD instead")
        white_noise = kwargs.pop("white_noise", None)          # pre-drawn by ensemble.ChainEnsemble (same stream order)
        delta = self.noisy_gradient(**kwargs)
        if white_noise is None:
            white_noise = self._get_sgmcmc_noise(**kwargs)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += epsilon * delta[var] + np.sqrt(2.0 * epsilon) * white_noise[var]
        return self.parameters

    def sample_sgld_cv(self, epsilon, centering_parameters, centering_gradient, **kwargs):
        if "preconditioner" in kwargs:
            raise ValueError("Use SGRLD instead")
        buffer_dicts = [self._random_subsequence_and_buffers(
            buffer_length=kwargs.get("buffer_length", 0),
            subsequence_length=kwargs.get("subsequence_length", -1), T=self._get_T(**kwargs))
            for _ in range(kwargs.get("minibatch_size", 1))]
        cur = self.noisy_gradient(buffer_dicts=buffer_dicts, **kwargs)
        cen = self.noisy_gradient(parameters=centering_parameters, buffer_dicts=buffer_dicts, **kwargs)
        delta = {var: centering_gradient[var] + cur[var] - cen[var] for var in cur}
        white_noise = self._get_sgmcmc_noise(**kwargs)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += epsilon * delta[var] + np.sqrt(2.0 * epsilon) * white_noise[var]
        return self.parameters

    def sample_sgrld(self, epsilon, preconditioner, **kwargs):
        """Riemannian Langevin step (sgmcmc_sampler.py:613-640)."""
        scale = 1.0 / self._get_T(**kwargs) if kwargs.get("is_scaled", True) else 1.0
        delta = self.noisy_gradient(preconditioner=preconditioner, **kwargs)
        white_noise = self._get_sgmcmc_noise(preconditioner=preconditioner, **kwargs)
        correction = preconditioner.correction_term(self.parameters, scale=scale)
        for var in self.parameters.var_dict:
            self.parameters.var_dict[var] += (epsilon * (delta[var] + correction[var]) +
                                              np.sqrt(2.0 * epsilon) * white_noise[var])
        return self.parameters

    def sample_gibbs(self):
        raise NotImplementedError()

    def project_parameters(self, **kwargs):
        self.parameters.project_parameters(**self.options, **kwargs)
        return self.parameters

    # ---- fit loops --------------------------------------------------------------------------------
    def fit(self, iter_type, num_iters, output_all=False, observations=None, init_parameters=None,
            tqdm=None, catch_interrupt=False, **kwargs):
        """sgmcmc_sampler.py:659-721."""
        if observations is not None:
            self.observations = observations
        if init_parameters is not None:
            self.parameters = init_parameters.copy()
        if self._device_loop_eligible(iter_type, kwargs):
            return self._fit_device(iter_type, num_iters, output_all=output_all, **kwargs)
        names, kws = self.get_iter_step(iter_type, tqdm=tqdm, **kwargs)
        parameters_list = [None] * (num_iters + 1)
        parameters_list[0] = self.parameters.copy()
        pbar = range(1, num_iters + 1)
        if tqdm is not None:
            pbar = tqdm(pbar)
            pbar.set_description("fit using {0} iters".format(iter_type))
        for it in pbar:
            try:
                for name, kw in zip(names, kws):
                    getattr(self, name)(**kw)
                if output_all:
                    parameters_list[it] = self.parameters.copy()
            except KeyboardInterrupt as e:
                if not catch_interrupt:
                    raise e
                logger.warning("Interrupt in fit:\n{0}\nStopping early after {1} iters".format(e, it))
                return parameters_list[:it] if output_all else self.parameters.copy()
        return parameters_list if output_all else self.parameters.copy()

    def _device_loop_eligible(self, iter_type, kwargs):
        """The whole loop runs on the device (device_loop.DeviceChains) when nothing in it needs the host: particle-filter
        gradients with device randoms, a plain SGLD / SGD / LGSSM-SGRLD step, default projection options."""
        from . import engine
        from .device_loop import _supported_options
        pfk = kwargs.get("pf_kwargs", {})
        if not kwargs.get("device_loop", engine.config.device_loop) or kwargs.get("kind") != "pf":
            return False
        if iter_type not in ("SGLD", "SGD", "SGRLD") or kwargs.get("steps_per_iteration", 1) != 1:
            return False
        if pfk.get("rng", engine.config.rng) != "philox" or kwargs.get("project_kwargs") or pfk.get("distributed"):
            return False
        if iter_type == "SGRLD":
            from .models.lgssm import LGSSMPreconditioner
            if not isinstance(kwargs.get("preconditioner"), LGSSMPreconditioner):
                return False
        return _supported_options(getattr(self, "options", {})) and getattr(self.parameters, "n", 1) == 1 \
            and getattr(self.parameters, "m", 1) == 1

    def _fit_device(self, iter_type, num_iters, output_all=False, **kwargs):
        """fit() with every iteration on the device (same arguments; windows, filter randoms and Langevin noise come
        from the device Philox streams instead of np.random)."""
        from .device_loop import DeviceChains
        pfk = dict(kwargs.get("pf_kwargs", {}))
        chains = DeviceChains([self], method=iter_type, epsilon=kwargs["epsilon"],
                              subsequence_length=kwargs["subsequence_length"], buffer_length=kwargs["buffer_length"],
                              minibatch_size=kwargs.get("minibatch_size", 1), num_sequences=kwargs.get("num_sequences"),
                              num_samples=kwargs.get("num_samples"), preconditioner=kwargs.get("preconditioner"),
                              trace_every=1 if output_all else 0, max_trace_rows=num_iters if output_all else 0, **pfk)
        chains.run(num_iters)
        chains.pull_parameters()
        if not output_all:
            return self.parameters.copy()
        tr = chains.trace()                                   # (num_iters + 1, 1, n_params)
        out = []
        for row in tr[:, 0, :]:
            p = self.parameters.copy()
            for i, k in enumerate(chains.slots):
                p.var_dict[k] = np.full_like(np.asarray(p.var_dict[k], dtype=float), row[i])
            out.append(p)
        return out

    def fit_timed(self, iter_type, max_time=60, min_save_time=1, observations=None, init_parameters=None,
                  tqdm=None, tqdm_iter=False, catch_interrupt=False, **kwargs):
        """Run for `max_time` seconds, saving parameters every `min_save_time` (sgmcmc_sampler.py:723-755)."""
        if observations is not None:
            self.observations = observations
        if init_parameters is not None:
            self.parameters = init_parameters.copy()
        names, kws = self.get_iter_step(iter_type, tqdm=tqdm, **kwargs)
        parameters_list, times = [self.parameters.copy()], [0.0]
        total, last = 0.0, time.time()
        while total <= max_time:
            try:
                for name, kw in zip(names, kws):
                    getattr(self, name)(**kw)
            except KeyboardInterrupt as e:
                if not catch_interrupt:
                    raise e
                break
            if time.time() - last > min_save_time:
                total += time.time() - last
                parameters_list.append(self.parameters.copy())
                times.append(total)
                last = time.time()
        return parameters_list, times

    def get_iter_step(self, iter_type, steps_per_iteration=1, **kwargs):
        """sgmcmc_sampler.py:896-947."""
        project_kwargs = kwargs.get("project_kwargs", {})
        if iter_type == "Gibbs":
            names, kws = ["sample_gibbs", "project_parameters"], [{}, project_kwargs]
        elif iter_type == "custom":
            names, kws = kwargs.get("iter_func_names"), kwargs.get("iter_func_kwargs")
        elif iter_type in ["SGD", "ADAGRAD", "SGLD", "SGRD", "SGRLD"]:
            grad_kwargs = dict(epsilon=kwargs["epsilon"], subsequence_length=kwargs["subsequence_length"],
                               buffer_length=kwargs["buffer_length"], minibatch_size=kwargs.get("minibatch_size", 1),
                               kind=kwargs.get("kind", "marginal"), num_samples=kwargs.get("num_samples", None),
                               **kwargs.get("pf_kwargs", {}))
            if "num_sequences" in kwargs:
                grad_kwargs["num_sequences"] = kwargs["num_sequences"]
            step = {"SGD": "step_sgd", "ADAGRAD": "step_adagrad", "SGLD": "sample_sgld",
                    "SGRD": "step_precondition_sgd", "SGRLD": "sample_sgrld"}[iter_type]
            if iter_type in ("SGRD", "SGRLD"):
                grad_kwargs["preconditioner"] = self._get_preconditioner(kwargs.get("preconditioner"))
            names, kws = [step, "project_parameters"], [grad_kwargs, project_kwargs]
        else:
            raise ValueError("Unrecognized iter_type {0}".format(iter_type))
        return names * steps_per_iteration, kws * steps_per_iteration

    def _get_preconditioner(self, preconditioner=None):
        if preconditioner is None:
            raise NotImplementedError("No Default Preconditioner for {}".format(self.name))
        return preconditioner

    # ---- prediction -------------------------------------------------------------------------------
    def predict(self, target="latent", distr=None, lag=None, return_distr=None, num_samples=None,
                kind="analytic", observations=None, parameters=None, **kwargs):
        """kind='pf', target='latent' -> smoothed marginal mean / covariance (sgmcmc_sampler.py:1047-1067)."""
        observations = self._get_observations(observations)
        if parameters is None:
            parameters = self.parameters
        if kind != "pf":
            if kind == "analytic":
                raise NotImplementedError(_NOT_PF.format(kind))
            raise ValueError("Unrecognized kind == '{0}'".format(kind))
        if return_distr is False:
            raise ValueError("return_distr must be True for kind = pf")
        if target != "latent":
            raise NotImplementedError("kind='pf' supports target='latent'")
        return self.message_helper.pf_latent_var_distr(lag=lag, observations=observations,
                                                       parameters=parameters, **kwargs)

    # ---- misc -------------------------------------------------------------------------------------
    @property
    def observations(self):
        return self._observations

    @observations.setter
    def observations(self, observations):
        self._check_observation_shape(observations)
        self._observations = observations

    def _check_observation_shape(self, observations):
        return

    def _get_observations(self, observations, check_shape=True):
        if observations is None:
            observations = self.observations
            if observations is None:
                raise ValueError("observations not specified")
        elif check_shape:
            self._check_observation_shape(observations)
        return observations

    def _get_T(self, **kwargs):
        T = kwargs.get("T")
        if T is None:
            T = self._get_observations(kwargs.get("observations")).shape[0]
        return T


class SeqSGMCMCSampler(object):
    """Mixin for a list of observation sequences (sgmcmc_sampler.py:1159-1283)."""

    def _get_T(self, **kwargs):
        T = kwargs.get("T")
        if T is None:
            observations = self._get_observations(kwargs.get("observations"))
            cache = getattr(self, "_T_cache", None)
            if cache is not None and cache[0] is observations and cache[1] == len(observations):
                return cache[2]
            T = int(np.sum([np.shape(o)[0] for o in observations]))
            self._T_cache = (observations, len(observations), T)
        return T

    def _check_observation_shape(self, observations):
        if observations is not None:
            for ii, observation in enumerate(observations):
                try:
                    super()._check_observation_shape(observations=observation)
                except ValueError as e:
                    raise ValueError("Error in observations[{0}] :\n{1}".format(ii, e))

    def _pick_sequences(self, observations, num_sequences):
        idx = np.arange(len(observations))
        if num_sequences != -1:
            idx = np.random.choice(idx, num_sequences, replace=False)
        return idx

    def noisy_loglikelihood(self, num_sequences=-1, observations=None, tqdm=None, **kwargs):
        observations = self._get_observations(observations)
        loglikelihood, S = 0.0, 0.0
        for k in self._pick_sequences(observations, num_sequences):
            S += observations[k].shape[0]
            loglikelihood += super().noisy_loglikelihood(observations=observations[k], **kwargs)
        if num_sequences != -1:
            loglikelihood *= self._get_T(**kwargs) / S
        return loglikelihood

    def _noisy_grad_loglikelihood(self, num_sequences=-1, subsequence_length=-1, minibatch_size=1,
                                  buffer_length=0, observations=None, buffer_dicts=None, kind="marginal",
                                  num_samples=None, parameters=None, **kwargs):
        """Sum over (sampled) sequences of per-sequence minibatch means, rescaled by T / S
        (sgmcmc_sampler.py:1249-1283).  Every (sequence, subsequence) pair is one work item of a single
        batched launch."""
        if kind != "pf":
            raise NotImplementedError(_NOT_PF.format(kind))
        from . import engine
        from .particle_filters.buffered_smoother import _draw_injected
        seqs = self.observations
        idx = self._pick_sequences(seqs, num_sequences)
        if kwargs.get("N", None) is None:
            kwargs["N"] = num_samples
        replay = kwargs.get("rng", engine.config.rng) == "injected" and "injected" not in kwargs
        windows, S, drawn = [], 0.0, []
        for k in idx:
            w = SGMCMCSampler._pf_windows(self, subsequence_length, minibatch_size, buffer_length,
                                          seqs[k], buffer_dicts)
            if replay:      # keep the reference's stream order: windows of sequence k, then its PF draws
                drawn.append(_draw_injected(int(kwargs["N"]), [x["observations"].shape[0] for x in w]))
            windows += w
            S += seqs[k].shape[0]
        if replay:
            kwargs["injected"] = _merge_injected(drawn)
            kwargs.setdefault("resample", "multinomial")
        grads, _ = self.message_helper.pf_gradient_estimate_batch(windows, self.parameters, **kwargs)
        noisy_grad = {var: np.zeros_like(value) for var, value in self.parameters.as_dict().items()}
        for g in grads:
            for var in noisy_grad:
                noisy_grad[var] += g[var] * 1.0 / minibatch_size
        for var in noisy_grad:
            if np.any(np.isnan(noisy_grad[var])):
                raise ValueError("NaNs in gradient of {0}".format(var))
        if num_sequences != -1:
            scale = self._get_T() / S
            noisy_grad = {var: noisy_grad[var] * scale for var in noisy_grad}
        return noisy_grad

    def _pf_plan(self, num_sequences=-1, subsequence_length=-1, minibatch_size=1, buffer_length=0, buffer_dicts=None,
                 **unused):
        """Seq form of SGMCMCSampler._pf_plan (sgmcmc_sampler.py:1249-1283): same numpy draws as
        _noisy_grad_loglikelihood (sequence choice, then the windows of each chosen sequence)."""
        seqs = self.observations
        idx = self._pick_sequences(seqs, num_sequences)
        windows, S = [], 0.0
        for k in idx:
            windows += SGMCMCSampler._pf_windows(self, subsequence_length, minibatch_size, buffer_length, seqs[k], buffer_dicts)
            S += seqs[k].shape[0]

        def finish(grads):
            noisy_grad = {var: np.zeros_like(value) for var, value in self.parameters.as_dict().items()}
            for g in grads:
                for var in noisy_grad:
                    noisy_grad[var] += g[var] * 1.0 / minibatch_size
            for var in noisy_grad:
                if np.any(np.isnan(noisy_grad[var])):
                    raise ValueError("NaNs in gradient of {0}".format(var))
            if num_sequences != -1:
                scale = self._get_T() / S
                noisy_grad = {var: noisy_grad[var] * scale for var in noisy_grad}
            return noisy_grad
        return windows, finish

    def predict(self, *args, **kwargs):
        raise NotImplementedError()
