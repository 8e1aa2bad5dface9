from ...particle_filters.kernels import LGSSMPriorKernel, LGSSMOptimalKernel, LGSSMHighDimOptimalKernel  # noqa: F401
