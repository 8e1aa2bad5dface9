"""sgmcmc_ssm_b200: B200-native buffered particle-filter gradient estimator behind the `sgmcmc_ssm`
Python API (LGSSM / SVM / GARCH).  Host code in Python/PyTorch, compute in hand-written sm_100a CUDA
kernels behind the C-ABI of include/sgmpf.h.  No CPU fallback."""
from .engine import config, set_seed, run_pf, PFItems, PackedItems  # noqa: F401

__version__ = "0.1.0"
