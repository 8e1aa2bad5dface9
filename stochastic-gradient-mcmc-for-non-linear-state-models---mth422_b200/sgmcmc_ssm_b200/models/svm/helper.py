import numpy as np

from ...helper import SGMCMCHelper
from ...particle_filters.kernels import SVMPriorKernel
from ...particle_filters.statistics import svm_complete_data_loglike_gradient  # noqa: F401


class SVMHelper(SGMCMCHelper):
    """sgmcmc_ssm/models/svm/helper.py:13-65."""
    _model = "svm"

    def __init__(self, n=1, m=1, forward_message=None, backward_message=None, **kwargs):
        self.n, self.m = n, m
        if forward_message is None:
            forward_message = {"log_constant": 0.0, "mean_precision": np.zeros(n), "precision": np.eye(n) / 10}
        self.default_forward_message = forward_message
        if backward_message is None:
            backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(n), "precision": np.zeros((n, n))}
        self.default_backward_message = backward_message

    def _get_kernel(self, kernel):
        if kernel is None:
            kernel = "prior"
        if kernel == "prior":
            return SVMPriorKernel()
        if kernel == "optimal":
            raise NotImplementedError("SVM optimal kernel not analytic")
        raise ValueError("Unrecoginized kernel = {0}".format(kernel))
