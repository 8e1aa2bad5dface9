import numpy as np

from ...sgmcmc_sampler import SGMCMCSampler, SeqSGMCMCSampler
from .parameters import GARCHPrior, GARCHParameters
from .helper import GARCHHelper


class GARCHSampler(SGMCMCSampler):
    """sgmcmc_ssm/models/garch/sampler.py:6-79."""

    def __init__(self, n=1, m=1, observations=None, prior=None, parameters=None, forward_message=None,
                 name="GARCHSampler", **kwargs):
        self.options = kwargs
        self.n, self.m, self.name = n, m, name
        self.setup(observations=observations, prior=prior, parameters=parameters, forward_message=forward_message)

    def setup(self, observations=None, prior=None, parameters=None, forward_message=None):
        self.observations = observations
        self.prior = GARCHPrior.generate_default_prior(n=self.n, m=self.m) if prior is None else prior
        if parameters is None:
            self.parameters = self.prior.sample_prior()
        else:
            if not isinstance(parameters, GARCHParameters):
                raise ValueError("parameters is not a GARCHParameter")
            self.parameters = parameters
        self.forward_message = forward_message
        self.backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(self.n),
                                 "precision": np.zeros((self.n, self.n))}
        self.message_helper = GARCHHelper(n=self.n, m=self.m, forward_message=forward_message,
                                          backward_message=self.backward_message)


class SeqGARCHSampler(SeqSGMCMCSampler, GARCHSampler):
    pass
