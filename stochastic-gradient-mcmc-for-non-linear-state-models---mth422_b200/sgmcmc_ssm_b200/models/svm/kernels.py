from ...particle_filters.kernels import SVMPriorKernel  # noqa: F401
