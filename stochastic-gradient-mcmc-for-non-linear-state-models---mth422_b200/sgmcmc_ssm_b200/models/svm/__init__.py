from .parameters import SVMParameters, SVMPrior, generate_svm_data  # noqa: F401
from .helper import SVMHelper  # noqa: F401
from .kernels import SVMPriorKernel  # noqa: F401
from .sampler import SVMSampler, SeqSVMSampler  # noqa: F401
