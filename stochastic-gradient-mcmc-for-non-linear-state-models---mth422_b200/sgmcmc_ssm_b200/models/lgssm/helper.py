import numpy as np

from ...helper import SGMCMCHelper
from ...particle_filters.kernels import LGSSMPriorKernel, LGSSMOptimalKernel, LGSSMHighDimOptimalKernel
from ...particle_filters.statistics import (lgssm_complete_data_loglike_gradient,  # noqa: F401
                                            gaussian_sufficient_statistics)  # noqa: F401


class LGSSMHelper(SGMCMCHelper):
    """PF members of sgmcmc_ssm/models/lgssm/helper.py (:30-50 ctor, :1089-1214 PF entry points)."""
    _model = "lgssm"

    def __init__(self, n=1, m=1, forward_message=None, backward_message=None, **kwargs):
        self.n, self.m = n, m
        if forward_message is None:
            forward_message = {"log_constant": 0.0, "mean_precision": np.zeros(n), "precision": np.eye(n) / 10}
        self.default_forward_message = forward_message
        if backward_message is None:
            backward_message = {"log_constant": 0.0, "mean_precision": np.zeros(n), "precision": np.zeros((n, n))}
        self.default_backward_message = backward_message

    def _get_kernel(self, kernel):
        if kernel is None:                     # lgssm/helper.py:1200-1214
            kernel = "optimal" if self.n * self.m == 1 else "highdim"
        if kernel == "prior":
            return LGSSMPriorKernel()
        if kernel == "optimal":
            return LGSSMOptimalKernel()
        if kernel == "highdim":
            raise NotImplementedError("LGSSMHighDimOptimalKernel (n > 1) is outside the CUDA path")
        raise ValueError("Unrecognized kernel = {0}".format(kernel))
