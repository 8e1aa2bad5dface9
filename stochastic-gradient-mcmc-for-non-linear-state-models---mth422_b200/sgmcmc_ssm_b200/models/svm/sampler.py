import numpy as np

from ...sgmcmc_sampler import SGMCMCSampler, SeqSGMCMCSampler
from .parameters import SVMPrior, SVMParameters
from .helper import SVMHelper


class SVMSampler(SGMCMCSampler):
    """sgmcmc_ssm/models/svm/sampler.py:6-81."""

    def __init__(self, n=1, m=1, observations=None, prior=None, parameters=None, forward_message=None,
                 name="SVMSampler", **kwargs):
        self.options = kwargs
        self.n, self.m, self.name = n, m, name
        self.setup(observations=observations, prior=prior, parameters=parameters, forward_message=forward_message)

    def setup(self, observations=None, prior=None, parameters=None, forward_message=None):
        self.observations = observations
        self.prior = SVMPrior.generate_default_prior(n=self.n, m=self.m) if prior is None else prior
        if parameters is None:
            self.parameters = self.prior.sample_prior().project_parameters()
        else:
            if not isinstance(parameters, SVMParameters):
                raise ValueError("parameters is not a SVMParameter")
            self.parameters = parameters
        if forward_message is None:
            forward_message = {"log_constant": 0.0, "mean_precision": np.zeros(self.n),
                               "precision": np.eye(self.n) / 10}
        self.forward_message = forward_message
        self.backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(self.n),
                                 "precision": np.zeros((self.n, self.n))}
        self.message_helper = SVMHelper(n=self.n, m=self.m, forward_message=forward_message,
                                        backward_message=self.backward_message)


class SeqSVMSampler(SeqSGMCMCSampler, SVMSampler):
    pass
