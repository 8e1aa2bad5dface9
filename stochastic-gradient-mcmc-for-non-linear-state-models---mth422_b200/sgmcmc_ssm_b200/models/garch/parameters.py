"""GARCH(1,1) observed in noise: sigma2_t = alpha + beta x_{t-1}^2 + gamma sigma2_{t-1},
x_t ~ N(0, sigma2_t), y_t ~ N(x_t, R).  API mirror of sgmcmc_ssm/models/garch/parameters.py:17-139."""
import numpy as np
from scipy.special import logit

from ...base_parameters import BaseParameters, BasePrior
from ...variables import GARCHVar, CovarianceVar

_G = GARCHVar()
_R = CovarianceVar("R", "m")


class GARCHParameters(BaseParameters):
    _variables = [_G, _R]

    def __str__(self):
        return "GARCHParameters:\nalpha:{0}, beta:{1}, gamma:{2}, tau:{3}\n".format(
            np.around(float(self.alpha[0]), 6), np.around(float(self.beta[0]), 6),
            np.around(float(self.gamma[0]), 6), np.around(float(np.ravel(self.tau)[0]), 6))

    @property
    def tau(self):
        return self.LRinv ** -1 if self.m == 1 else np.linalg.inv(self.LRinv.T)

    @staticmethod
    def convert_alpha_beta_gamma(alpha, beta, gamma):
        """(alpha, beta, gamma) -> (log_mu, logit_phi, logit_lambduh)  (garch/parameters.py:45-60)."""
        if alpha <= 0 or beta <= 0 or gamma <= 0:
            raise ValueError("Cannot have alpha, beta, or gamma <= 0")
        if beta + gamma >= 1:
            raise ValueError("Cannot have beta + gamma >- 1")
        return np.log(alpha / (1 - beta - gamma)), logit(beta + gamma), logit(beta / (beta + gamma))


class GARCHPrior(BasePrior):
    _Parameters = GARCHParameters
    _variables = [_G, _R]


def generate_garch_data(T, parameters, initial_message=None, tqdm=None):
    """Synthetic GARCH series (garch/parameters.py:74-139); reference draw order."""
    n = m = 1
    alpha, beta, gamma, R = parameters.alpha, parameters.beta, parameters.gamma, parameters.R
    if initial_message is None:
        initial_message = {"log_constant": 0.0, "mean_precision": np.zeros(n),
                           "precision": np.atleast_2d((1 - beta - gamma) / alpha)}
    latent_vars, sigma2s, obs_vars = np.zeros((T, n)), np.zeros(T), np.zeros((T, m))
    prev = np.random.multivariate_normal(
        mean=np.linalg.solve(initial_message["precision"], initial_message["mean_precision"]),
        cov=np.linalg.inv(initial_message["precision"]))
    sigma2_prev = 0
    for t in range(T):
        sigma2s[t] = float(np.ravel(alpha + beta * prev ** 2 + gamma * sigma2_prev)[0])
        latent_vars[t] = np.random.multivariate_normal(mean=np.zeros(1), cov=np.array([[sigma2s[t]]]))
        obs_vars[t] = np.random.multivariate_normal(mean=latent_vars[t], cov=R)
        prev = latent_vars[t]
        sigma2_prev = sigma2s[t]
    return dict(observations=obs_vars, latent_vars=latent_vars, sigma2s=sigma2s, parameters=parameters,
                initial_message=initial_message)
