"""Drop-in alias: put this directory on PYTHONPATH and `import sgmcmc_ssm` resolves to the B200
implementation (`sgmcmc_ssm_b200`) under the reference's module paths, e.g.
    from sgmcmc_ssm.models.svm import SVMSampler, SVMParameters, generate_svm_data
    from sgmcmc_ssm.particle_filters.buffered_smoother import buffered_pf_wrapper
"""
import importlib
import os
import sys

_here = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _here not in sys.path:
    sys.path.insert(0, _here)

import sgmcmc_ssm_b200 as _impl  # noqa: E402

_SUBMODULES = [
    "_utils", "base_parameters", "variables", "sgmcmc_sampler", "helper", "engine", "trace_metric_functions", "metric_functions",
    "ensemble",
    "particle_filters", "particle_filters.buffered_smoother", "particle_filters.pf", "particle_filters.kernels",
    "models", "models.svm", "models.svm.parameters", "models.svm.helper", "models.svm.kernels", "models.svm.sampler",
    "models.lgssm", "models.lgssm.parameters", "models.lgssm.helper", "models.lgssm.kernels", "models.lgssm.sampler",
    "models.garch", "models.garch.parameters", "models.garch.helper", "models.garch.kernels", "models.garch.sampler",
]
for _name in _SUBMODULES:
    sys.modules[__name__ + "." + _name] = importlib.import_module("sgmcmc_ssm_b200." + _name)

config = _impl.config
set_seed = _impl.set_seed
__version__ = _impl.__version__
