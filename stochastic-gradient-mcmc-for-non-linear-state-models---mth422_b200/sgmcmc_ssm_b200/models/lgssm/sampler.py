import numpy as np

from ...sgmcmc_sampler import SGMCMCSampler, SeqSGMCMCSampler
from .parameters import LGSSMPrior, LGSSMPreconditioner
from .helper import LGSSMHelper


class LGSSMSampler(SGMCMCSampler):
    """sgmcmc_ssm/models/lgssm/sampler.py:6-100 (PF path)."""

    def __init__(self, n=1, m=1, observations=None, prior=None, parameters=None, forward_message=None,
                 backward_message=None, name="LGSSMSampler", **kwargs):
        self.options = kwargs
        self.n, self.m, self.name = n, m, name
        self.setup(observations=observations, prior=prior, parameters=parameters,
                   forward_message=forward_message, backward_message=backward_message)

    def setup(self, observations=None, prior=None, parameters=None, forward_message=None, backward_message=None):
        self.observations = observations
        self.prior = LGSSMPrior.generate_default_prior(n=self.n, m=self.m) if prior is None else prior
        self.parameters = self.prior.sample_prior() if parameters is None else parameters
        if forward_message is None:
            forward_message = {"log_constant": 0.0, "mean_precision": np.zeros(self.n),
                               "precision": np.eye(self.n) / 10}
        self.forward_message = forward_message
        if backward_message is None:
            backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(self.n),
                                "precision": np.zeros((self.n, self.n))}
        self.backward_message = backward_message
        self.message_helper = LGSSMHelper(n=self.n, m=self.m, forward_message=forward_message,
                                          backward_message=backward_message)

    def _check_observation_shape(self, observations):
        if observations is None:
            return
        if np.shape(observations)[1] != self.m:
            raise ValueError("observations second dimension does not match m")

    def _get_preconditioner(self, preconditioner=None):
        return LGSSMPreconditioner() if preconditioner is None else preconditioner


class SeqLGSSMSampler(SeqSGMCMCSampler, LGSSMSampler):
    pass
