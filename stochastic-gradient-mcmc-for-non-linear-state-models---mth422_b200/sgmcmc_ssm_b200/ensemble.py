"""Independent SG-MCMC chains advanced in lock-step with ONE batched device launch per iteration.

The reference runs independent chains as separate samplers / processes (one per experiment id,
nonlinear_ssm_pf_experiment_scripts/*/driver.py:466-496; one per seed in demo/exchange_rate/*.py) -- each step of each
chain is its own Python loop over tiny particle filters.  Here every chain's (sequence, subsequence) work items of
an iteration -- each with its OWN parameters theta_c and prior -- are packed into one C-ABI call (items are
independent); the per-chain numpy streams (window draws, SGLD noise) stay exactly those of a separate
process seeded with that chain's seed.

Multi-GPU: chains shard over ranks with no collective at all -- build the ensemble from
`samplers[lo:hi]` with `lo, hi = parallel.shard_bounds(n_chains)`; gather parameter traces at the end.
"""
import contextlib

import numpy as np

from . import engine
from .particle_filters import statistics as S
from .particle_filters.buffered_smoother import batched_pf


class ChainEnsemble(object):
    def __init__(self, samplers, seeds):
        self.samplers = list(samplers)
        if len(self.samplers) == 0:
            raise ValueError("empty ensemble")
        if len(seeds) != len(self.samplers):
            raise ValueError("one seed per chain")
        models = {s.message_helper._model for s in self.samplers}
        if len(models) != 1:
            raise ValueError("all chains of an ensemble must share the model")
        self.model = models.pop()
        self._states = [np.random.RandomState(int(seed)).get_state() for seed in seeds]

    @contextlib.contextmanager
    def _stream(self, c):
        """Run a block on chain c's own legacy numpy stream (what a separate process seeded with its seed would see)."""
        outer = np.random.get_state()
        np.random.set_state(self._states[c])
        try:
            yield
        finally:
            self._states[c] = np.random.get_state()
            np.random.set_state(outer)

    def __len__(self):
        return len(self.samplers)

    def noisy_grad_loglikelihoods(self, kind="pf", pf="poyiadjis_N", N=None, num_samples=None, kernel=None,
                                  forward_message=None, predraw_noise=False, **kwargs):
        """One dict of likelihood-gradient estimates per chain (what SGMCMCSampler._noisy_grad_loglikelihood
        returns), all chains in a single launch."""
        if kind != "pf":
            raise NotImplementedError("ChainEnsemble batches the particle-filter path only (kind='pf')")
        N = num_samples if N is None else N
        if N is None:
            raise TypeError("N (number of particles) must be given for kind='pf'")
        plans, parts = [], []
        self._noise = [None] * len(self.samplers)
        for c, s in enumerate(self.samplers):
            with self._stream(c):
                windows, finish = s._pf_plan(**kwargs)
                if predraw_noise:          # the SGLD noise does not depend on the gradient: draw it now, in the
                    self._noise[c] = s._get_sgmcmc_noise(**kwargs)      # chain's stream order (windows, then noise)
            plans.append((len(windows), finish))
            parts.append(s.message_helper.make_items(windows, s.parameters, forward_message))
        K = self.samplers[0].message_helper._get_kernel(kernel)
        for s in self.samplers:
            s.message_helper._get_kernel(kernel).set_parameters(s.parameters)      # |A| > 1 etc. raise here
        res = batched_pf(pf, K.model, K.kernel, engine.PackedItems.concat(parts), N, stat_kind="score", **kwargs)
        keys = S.MODEL_SPECS[self.model]["grad_keys"]
        out, lo = [], 0
        for n_items, finish in plans:
            grads = [{k: g[i] for i, k in enumerate(keys)} for g in res.grad[lo:lo + n_items]]
            out.append(finish(grads))
            lo += n_items
        self.last_result = res
        return out

    def _step(self, method, epsilon, project=True, **kwargs):
        grads = self.noisy_grad_loglikelihoods(**kwargs)
        for c, s in enumerate(self.samplers):
            with self._stream(c):
                getattr(s, method)(epsilon=epsilon, noisy_grad_loglike=grads[c], **kwargs)
                if project:
                    s.project_parameters()
        return [s.parameters for s in self.samplers]

    def sample_sgld(self, epsilon, project=True, **kwargs):
        """SGLD step + project_parameters of every chain (sgmcmc_sampler.py:549-567, 650-656).  The noise is drawn
        together with the windows (it does not depend on the gradient), so each chain's stream is entered once."""
        grads = self.noisy_grad_loglikelihoods(predraw_noise=True, **kwargs)
        for c, s in enumerate(self.samplers):
            s.sample_sgld(epsilon=epsilon, noisy_grad_loglike=grads[c], white_noise=self._noise[c], **kwargs)
            if project:
                s.project_parameters()                  # deterministic: no stream needed
        return [s.parameters for s in self.samplers]

    def step_sgd(self, epsilon, **kwargs):
        return self._step("step_sgd", epsilon, **kwargs)

    def step_adagrad(self, epsilon, **kwargs):
        return self._step("step_adagrad", epsilon, **kwargs)

    def fit(self, iter_type, num_iters, epsilon, output_all=False, trace_every=None, **kwargs):
        """`num_iters` iterations of every chain on the device with no host round trip (device_loop.DeviceChains: windows,
        particle filters, prior gradient, Langevin noise, update and projection in one launch sequence per iteration, all
        chains batched).  Randomness comes from the device Philox streams (keyed by chain index, so sharding the
        chains over GPUs does not change a chain's stream).  Returns the parameter trace (rows, chains, n_params) when
        output_all, else the list of final Parameters."""
        from .device_loop import DeviceChains
        pfk = dict(kwargs.pop("pf_kwargs", {}))
        pfk.update(kwargs)
        every = (trace_every or 1) if output_all else 0
        chains = DeviceChains(self.samplers, method=iter_type, epsilon=epsilon, trace_every=every,
                              max_trace_rows=(num_iters // every if every else 0), **pfk)
        chains.run(num_iters)
        params = chains.pull_parameters()
        self.device_chains = chains
        return chains.trace() if output_all else params
