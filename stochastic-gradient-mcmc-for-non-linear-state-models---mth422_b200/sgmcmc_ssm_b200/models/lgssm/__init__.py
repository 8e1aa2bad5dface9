from .parameters import LGSSMParameters, LGSSMPrior, LGSSMPreconditioner, generate_lgssm_data  # noqa: F401
from .helper import LGSSMHelper  # noqa: F401
from .kernels import LGSSMPriorKernel, LGSSMOptimalKernel, LGSSMHighDimOptimalKernel  # noqa: F401
from .sampler import LGSSMSampler, SeqLGSSMSampler  # noqa: F401
