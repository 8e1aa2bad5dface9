"""Metric functions for full parameter traces -- drop-in for sgmcmc_ssm/trace_metric_functions.py: `IMQ_KSD`
(:20-81) and `compute_KSD` (:83-117), with the all-pairs sum on the GPU (csrc/ksd_kernel.cuh through the C-ABI
`sgm_ksd_imq`).  SURVEY 8(f4)."""
import ctypes
import logging

import numpy as np
import torch

from . import _native as nat
from . import engine

logger = logging.getLogger(name=__name__)


def IMQ_KSD(x, gradlogp, c=1, beta=0.5, max_block_size=1000, tqdm_out=None):
    """Inverse-multiquadric kernel Stein discrepancy, IMQ(x, y) = (c^2 + |x - y|^2)^-beta
    (trace_metric_functions.py:20-81).  x, gradlogp: (num_points, d) arrays, d <= 8.  `max_block_size` / `tqdm_out`
    are accepted for signature compatibility (the reference blocks its numpy temporaries with them)."""
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    gradlogp = np.ascontiguousarray(np.asarray(gradlogp, dtype=np.float64))
    if x.shape != gradlogp.shape:
        raise ValueError("x and gradlogp dimensions do not match")
    if x.ndim != 2:
        raise ValueError("x must be num_points by d")
    K, d = x.shape
    lib = nat.load()
    device = engine._device()
    engine._state(device)                                   # device check (B200 only, no fallback)
    with torch.cuda.device(device):
        xd, gd = torch.from_numpy(x).to(device), torch.from_numpy(gradlogp).to(device)
        partial = torch.empty((K + 255) // 256, dtype=torch.float64, device=device)
        stream = torch.cuda.current_stream(device)
        nat.check(lib.sgm_ksd_imq(xd.data_ptr(), gd.data_ptr(), K, d, float(c), float(beta), partial.data_ptr(),
                                  ctypes.c_void_p(stream.cuda_stream)))
        total = float(partial.sum().item())
    return np.sqrt(total) / K


def compute_KSD(param_list, grad_list, variables=None, **kwargs):
    """trace_metric_functions.py:83-117: one IMQ_KSD per variable of a list of Parameters / gradient vectors."""
    res = {}
    if variables is not None:
        for ii, var in enumerate(variables):
            if hasattr(param_list[0], var):
                x = np.array([np.asarray(getattr(parameters, var)).flatten() for parameters in param_list])
                gradlogp = np.array([grad[ii] for grad in grad_list])
                if len(x.shape) == 1:
                    x = np.reshape(x, (-1, 1))
                if len(gradlogp.shape) == 1:
                    gradlogp = np.reshape(gradlogp, (-1, 1))
                res[var] = IMQ_KSD(x, gradlogp, **kwargs)
            else:
                logger.warning("Did not find {0} in parameters".format(var))
    return res
