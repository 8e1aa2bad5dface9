"""Proposal-kernel tags.  In the reference these classes hold the numpy proposal / reweight code
(particle_filters/kernels.py:9-138, models/*/kernels.py); here the arithmetic lives in the CUDA
functors (csrc/models.cuh) and the Python objects only select (model, kernel) for the C-ABI and keep
the reference's validation behaviour."""
import numpy as np


class Kernel(object):
    model = None
    kernel = None

    def __init__(self, **kwargs):
        self.parameters = kwargs.get("parameters", None)
        self.y_next = kwargs.get("y_next", None)

    def set_parameters(self, parameters):
        self.parameters = parameters

    def set_y_next(self, y_next):
        self.y_next = y_next


class LatentGaussianKernel(Kernel):
    pass


class SVMPriorKernel(LatentGaussianKernel):
    model, kernel = "svm", "prior"

    def set_parameters(self, parameters):
        self.parameters = parameters
        if np.abs(parameters.A) > 1:          # models/svm/kernels.py:8-10
            raise ValueError("Current AR parameter is |A| = {0} > 1".format(np.abs(parameters.A)) +
                             "\nTry calling project_parameters?")


class LGSSMPriorKernel(LatentGaussianKernel):
    model, kernel = "lgssm", "prior"


class LGSSMOptimalKernel(LatentGaussianKernel):
    model, kernel = "lgssm", "optimal"


class LGSSMHighDimOptimalKernel(LatentGaussianKernel):
    model, kernel = "lgssm", "highdim"       # n > 1: out of scope of the CUDA path


class GARCHPriorKernel(Kernel):
    model, kernel = "garch", "prior"


class GARCHOptimalKernel(Kernel):
    model, kernel = "garch", "optimal"
