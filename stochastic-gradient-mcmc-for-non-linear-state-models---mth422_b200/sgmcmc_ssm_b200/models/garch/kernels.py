from ...particle_filters.kernels import GARCHPriorKernel, GARCHOptimalKernel  # noqa: F401
