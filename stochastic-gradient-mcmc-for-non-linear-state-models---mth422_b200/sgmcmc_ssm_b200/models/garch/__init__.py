from .parameters import GARCHParameters, GARCHPrior, generate_garch_data  # noqa: F401
from .helper import GARCHHelper  # noqa: F401
from .kernels import GARCHPriorKernel, GARCHOptimalKernel  # noqa: F401
from .sampler import GARCHSampler, SeqGARCHSampler  # noqa: F401
