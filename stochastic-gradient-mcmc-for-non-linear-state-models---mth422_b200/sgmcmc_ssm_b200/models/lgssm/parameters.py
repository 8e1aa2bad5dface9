"""Linear-Gaussian SSM: x_t = A x_{t-1} + N(0, Q),  y_t = C x_t + N(0, R).
API mirror of sgmcmc_ssm/models/lgssm/parameters.py:18-129."""
import numpy as np

from ...base_parameters import BaseParameters, BasePrior, BasePreconditioner
from ...variables import SquareMatrixVar, RectMatrixVar, CovarianceVar
from ..._utils import var_stationary_precision

_A = SquareMatrixVar("A", "n", var_row_name="Q")
_C = RectMatrixVar("C", ("m", "n"), var_row_name="R")
_Q = CovarianceVar("Q", "n", matrix_name="A")
_R = CovarianceVar("R", "m", matrix_name="C")


class LGSSMParameters(BaseParameters):
    _variables = [_A, _C, _Q, _R]

    def __str__(self):
        return "LGSSMParameters:\nA:\n{0}\nC:\n{1}\nQ:\n{2}\nR:\n{3}".format(self.A, self.C, self.Q, self.R)

    def project_parameters(self, **kwargs):
        if "C" not in kwargs:                  # lgssm/parameters.py:39-42
            kwargs["C"] = dict(fixed_eye=True)
        return super().project_parameters(**kwargs)


class LGSSMPrior(BasePrior):
    _Parameters = LGSSMParameters
    _variables = [_Q, _R, _A, _C]


class LGSSMPreconditioner(BasePreconditioner):
    """lgssm/parameters.py:58-67."""
    _variables = [_A, _C, _Q, _R]


def generate_lgssm_data(T, parameters, initial_message=None, tqdm=None):
    """Synthetic LGSSM series (lgssm/parameters.py:69-129); reference draw order."""
    m, n = np.shape(parameters.C)
    A, C, Q, R = parameters.A, parameters.C, parameters.Q, parameters.R
    if initial_message is None:
        initial_message = {"log_constant": 0.0, "mean_precision": np.zeros(n),
                           "precision": var_stationary_precision(parameters.Qinv, parameters.A, 10)}
    latent_vars, obs_vars = np.zeros((T, n)), np.zeros((T, m))
    prev = np.random.multivariate_normal(
        mean=np.linalg.solve(initial_message["precision"], initial_message["mean_precision"]),
        cov=np.linalg.inv(initial_message["precision"]))
    for t in range(T):
        latent_vars[t] = np.random.multivariate_normal(mean=np.dot(A, prev), cov=Q)
        obs_vars[t] = np.random.multivariate_normal(mean=np.dot(C, latent_vars[t]), cov=R)
        prev = latent_vars[t]
    return dict(observations=obs_vars, latent_vars=latent_vars, parameters=parameters,
                initial_message=initial_message)
