import numpy as np

from ...helper import SGMCMCHelper
from ...particle_filters.kernels import GARCHPriorKernel, GARCHOptimalKernel
from ...particle_filters.statistics import (garch_complete_data_loglike_gradient,  # noqa: F401
                                            garch_sufficient_statistics)  # noqa: F401


class GARCHHelper(SGMCMCHelper):
    """PF members of sgmcmc_ssm/models/garch/helper.py:12-57, :324-332."""
    _model = "garch"

    def __init__(self, n=1, m=1, forward_message=None, backward_message=None, **kwargs):
        self.n, self.m = n, m
        self.default_forward_message = forward_message
        if backward_message is None:
            backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(n), "precision": np.zeros((n, n))}
        self.default_backward_message = backward_message

    def _get_kernel(self, kernel):
        if kernel is None:
            kernel = "optimal"
        if kernel == "prior":
            return GARCHPriorKernel()
        if kernel == "optimal":
            return GARCHOptimalKernel()
        raise ValueError("Unrecoginized kernel = {0}".format(kernel))

    def _prior_moments(self, forward_message, parameters):
        # garch/helper.py:324-332: the *argument* decides (the helper's default message is not consulted)
        if forward_message is None:
            return 0.0, float(np.ravel(parameters.alpha / (1 - parameters.beta - parameters.gamma))[0])
        prior_var = np.linalg.inv(np.atleast_2d(forward_message["precision"]))
        prior_mean = np.linalg.solve(prior_var, np.atleast_1d(forward_message["mean_precision"]))
        return float(prior_mean[0]), float(prior_var[0, 0])

    def _get_prior_x(self, forward_message, parameters):
        return self._prior_moments(forward_message, parameters)
