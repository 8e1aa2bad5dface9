"""Stochastic-volatility model: x_t = A x_{t-1} + N(0, Q),  y_t ~ N(0, R exp(x_t)).
API mirror of sgmcmc_ssm/models/svm/parameters.py:20-135."""
import numpy as np

from ...base_parameters import BaseParameters, BasePrior
from ...variables import SquareMatrixVar, CovarianceVar
from ..._utils import var_stationary_precision


class SVMParameters(BaseParameters):
    _variables = [SquareMatrixVar("A", "n", var_row_name="Q"), CovarianceVar("Q", "n", matrix_name="A"),
                  CovarianceVar("R", "m")]

    def __str__(self):
        if self.n == 1:
            return "SVMParameters:\nA:{0}, Q:{1}, R:{2}\n".format(self.A[0, 0], self.Q[0, 0], self.R[0, 0])
        return "SVMParameters:\nA:\n{0}\nQ:\n{1}\nR:\n{2}".format(self.A, self.Q, self.R)

    @property
    def phi(self):
        return self.A

    @property
    def sigma(self):
        return self.LQinv ** -1 if self.n == 1 else np.linalg.inv(self.LQinv.T)

    @property
    def tau(self):
        return self.LRinv ** -1 if self.m == 1 else np.linalg.inv(self.LRinv.T)


class SVMPrior(BasePrior):
    _Parameters = SVMParameters
    _variables = [CovarianceVar("Q", "n", matrix_name="A"), CovarianceVar("R", "m"),
                  SquareMatrixVar("A", "n", var_row_name="Q")]


def generate_svm_data(T, parameters, initial_message=None, tqdm=None):
    """Synthetic SVM series (svm/parameters.py:75-135); same draw order from the global numpy stream
    (one multivariate_normal per latent / observation) so a fixed seed gives the reference's data."""
    n, m = np.shape(parameters.A)[0], np.shape(parameters.R)[0]
    A, Q, R = parameters.A, parameters.Q, parameters.R
    if initial_message is None:
        initial_message = {"log_constant": 0.0, "mean_precision": np.zeros(n),
                           "precision": var_stationary_precision(parameters.Qinv, parameters.A, 10)}
    latent_vars, obs_vars = np.zeros((T, n)), np.zeros((T, m))
    prev = np.random.multivariate_normal(
        mean=np.linalg.solve(initial_message["precision"], initial_message["mean_precision"]),
        cov=np.linalg.inv(initial_message["precision"]))
    for t in range(T):
        latent_vars[t] = np.random.multivariate_normal(mean=np.dot(A, prev), cov=Q)
        obs_vars[t] = np.random.multivariate_normal(mean=np.zeros(1), cov=np.exp(latent_vars[t]) * R)
        prev = latent_vars[t]
    return dict(observations=obs_vars, latent_vars=latent_vars, parameters=parameters,
                initial_message=initial_message)
