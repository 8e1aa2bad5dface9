"""Model variables: each class bundles the parameter storage / reparametrisation, the prior and the
(optional) preconditioner of ONE variable.  Behavioural references (sgmcmc_ssm/variables/):
  covariance.py:19-156 (param), :158-284 (prior), :286-317 (precond)   -> CovarianceVar
  matrices.py:448-501, :503-630, :632-656                              -> SquareMatrixVar
  matrices.py:907-965, :967-1092, :1094-1125                           -> RectMatrixVar
  garch_var.py:21-91, :93-189                                          -> GARCHVar
"""
import logging

import numpy as np
import scipy.stats
from scipy.special import expit, logit

from ._utils import (array_wishart_rvs, matrix_normal_logpdf, pos_def_mat_inv, tril_indices_from, tril_vector_to_mat,
                     varp_stability_projection)

logger = logging.getLogger(name=__name__)


def _dim_prop(name):
    return property(lambda self: self.dim[name])


def _value_prop(name):
    def fset(self, value):
        self.var_dict[name] = value
    return property(lambda self: self.var_dict[name], fset)


class CovarianceVar(object):
    """Covariance Q stored as the packed Cholesky factor of its precision, `L<Q>inv_vec`."""

    def __init__(self, name="Q", dim_name="n", matrix_name=None):
        self.name, self.dim_name, self.matrix_name = name, dim_name, matrix_name
        self.vec, self.chol, self.inv = "L{0}inv_vec".format(name), "L{0}inv".format(name), "{0}inv".format(name)
        self.scale, self.df = "scale_{0}inv".format(name), "df_{0}inv".format(name)

    # -- parameter role
    def init_param(self, param, **kw):
        if self.vec in kw:
            vec = np.array(kw[self.vec]).astype(float)
            n = int(np.sqrt(len(vec) * 2))
        elif self.chol in kw or self.name in kw:
            if self.chol in kw:
                L = np.array(kw[self.chol]).astype(float)
            else:
                L = np.linalg.cholesky(np.linalg.inv(np.array(kw[self.name]).astype(float)))
            n, n2 = np.shape(L)
            if n != n2:
                raise ValueError("{0} must be square matrix".format(self.chol))
            vec = L[tril_indices_from(L)]
        else:
            raise ValueError("{0} not provided".format(self.chol))
        param.var_dict[self.vec] = vec
        param._set_check_dim(**{self.dim_name: n})

    def properties(self):
        vec, chol, inv = self.vec, self.chol, self.inv

        def get_chol(p):
            return tril_vector_to_mat(p.var_dict[vec])

        def set_chol(p, value):
            p.var_dict[vec] = value[tril_indices_from(value)]

        def get_inv(p):                       # covariance.py:141-146: L L^T + 1e-16 I
            L = get_chol(p)
            if L.size == 1:                   # 1 x 1: the same two roundings as the matrix expression
                return L * L + 1e-16
            return L.dot(L.T) + 1e-16 * np.eye(L.shape[0])

        def get_cov(p):
            Qinv = get_inv(p)
            return Qinv ** -1 if np.size(Qinv) == 1 else pos_def_mat_inv(Qinv)

        return {vec: _value_prop(vec), chol: property(get_chol, set_chol), inv: property(get_inv),
                self.name: property(get_cov), self.dim_name: _dim_prop(self.dim_name)}

    def project(self, param, **kwargs):
        opts = kwargs.get(self.name, {})
        if opts.get("fixed") is not None:
            param.var_dict[self.vec] = opts["fixed"].copy()
        if opts.get("thresh", True):
            L = tril_vector_to_mat(param.var_dict[self.vec])
            if np.any(np.diag(L) < 0.0):      # reflect: covariance.py:67-79
                logger.info("Reflecting {0}: {1} < 0.0".format(self.chol, L))
                L = np.linalg.cholesky(np.dot(L, L.T) + np.eye(L.shape[0]) * 1e-16)
            param.var_dict[self.vec] = L[tril_indices_from(L)]

    def flatten(self, var_dict):
        return [np.atleast_1d(var_dict[self.vec])]

    def unflatten(self, out, vector, idx, **dim):
        n = dim[self.dim_name]
        k = (n + 1) * n // 2
        out[self.vec] = np.array(vector[idx:idx + k], dtype=float)
        return idx + k

    # -- prior role (Wishart on the precision)
    def init_prior(self, prior, **kw):
        if self.scale not in kw:
            raise ValueError("{0} must be provided".format(self.scale))
        if self.df not in kw:
            raise ValueError("{0} must be provided".format(self.df))
        n, n2 = np.shape(kw[self.scale])
        if n != n2:
            raise ValueError("{0} must be square".format(self.scale))
        prior._set_check_dim(**{self.dim_name: n})
        prior.hyperparams[self.scale] = kw[self.scale]
        prior.hyperparams[self.df] = kw[self.df]

    def sample_prior(self, prior, var_dict):
        Qinv = array_wishart_rvs(df=prior.hyperparams[self.df], scale=prior.hyperparams[self.scale])
        L = np.linalg.cholesky(Qinv)
        var_dict[self.vec] = L[tril_indices_from(L)]

    def logprior(self, prior, parameters):
        return scipy.stats.wishart.logpdf(getattr(parameters, self.inv), df=prior.hyperparams[self.df],
                                          scale=prior.hyperparams[self.scale])

    def grad_logprior(self, prior, grad, parameters):
        L = getattr(parameters, self.chol)
        if L.size == 1:
            # 1 x 1: LAPACK's inv / solve reduce to one division each (bitwise equal, checked on 2e5 random inputs)
            g = (prior.hyperparams[self.df] - 2) * (1.0 / L) - L / np.reshape(prior.hyperparams[self.scale], (1, 1))
        else:
            g = ((prior.hyperparams[self.df] - L.shape[0] - 1) * np.linalg.inv(L.T)
                 - np.linalg.solve(prior.hyperparams[self.scale], L))
        grad[self.vec] = g[tril_indices_from(g)]

    def _hyper(self, kw, Qinv, var):
        df = np.shape(Qinv)[-1] + 1.0 + var ** -1
        kw[self.scale] = Qinv / df
        kw[self.df] = df

    def prior_kwargs_from(self, kw, parameters, from_mean=False, var=1.0):
        Qinv = getattr(parameters, self.inv) if from_mean else np.eye(getattr(parameters, self.chol).shape[0])
        self._hyper(kw, Qinv, var)

    def default_prior_kwargs(self, kw, var=100.0, **dims):
        self._hyper(kw, np.eye(dims[self.dim_name]), var)

    # -- preconditioner role
    def precondition(self, out, grad, parameters):
        Qinv = getattr(parameters, self.inv)
        G = np.zeros(Qinv.shape)
        G[tril_indices_from(G)] = grad[self.vec]
        P = np.dot(0.5 * Qinv, G)
        out[self.vec] = P[tril_indices_from(P)]

    def precondition_noise(self, out, parameters):
        L = tril_vector_to_mat(parameters.var_dict[self.vec])
        Z = np.dot(np.sqrt(0.5) * L, np.random.normal(loc=0, size=L.shape))
        out[self.vec] = Z[tril_indices_from(Z)]

    def correction_term(self, out, parameters):
        vec = parameters.var_dict[self.vec]
        n = int(np.sqrt(len(vec) * 2))
        out[self.vec] = 0.5 * (n + 1) * vec


class _MatrixVar(object):
    """Shared by SquareMatrixVar / RectMatrixVar: matrix-normal prior with row covariance
    `var_row_name` (matrices.py:503-630, :967-1092)."""

    def _setup(self, name, var_row_name):
        self.name, self.var_row_name = name, var_row_name
        self.mean, self.var_col = "mean_{0}".format(name), "var_col_{0}".format(name)
        self.row_vec = "L{0}inv_vec".format(var_row_name)

    def flatten(self, var_dict):
        return [var_dict[self.name].flatten()]

    def _row_precision(self, prior, var_dict, nrow):
        if self.var_row_name is None:
            return np.eye(nrow)
        if self.row_vec not in var_dict:
            raise ValueError("Missing {0}: {1} must be earlier in the prior's variable list".format(
                self.row_vec, self.var_row_name))
        L = tril_vector_to_mat(var_dict[self.row_vec])
        return L.dot(L.T) + 1e-9 * np.eye(nrow)

    def sample_prior(self, prior, var_dict):
        mean = prior.hyperparams[self.mean]
        Qinv = self._row_precision(prior, var_dict, mean.shape[0])
        draw = scipy.stats.matrix_normal(mean=mean, rowcov=pos_def_mat_inv(Qinv),
                                         colcov=np.diag(prior.hyperparams[self.var_col])).rvs()
        var_dict[self.name] = np.reshape(draw, mean.shape)

    def logprior(self, prior, parameters):
        mean, var_col = prior.hyperparams[self.mean], prior.hyperparams[self.var_col]
        if self.var_row_name is not None:
            L = tril_vector_to_mat(parameters.var_dict[self.row_vec])
        else:
            L = np.eye(mean.shape[0])
        return matrix_normal_logpdf(parameters.var_dict[self.name], mean=mean, Lrowprec=L,
                                    Lcolprec=np.diag(var_col ** -0.5))

    def grad_logprior(self, prior, grad, parameters):
        mean, var_col = prior.hyperparams[self.mean], prior.hyperparams[self.var_col]
        A = getattr(parameters, self.name)
        Qinv = getattr(parameters, "{0}inv".format(self.var_row_name)) if self.var_row_name else np.eye(A.shape[0])
        grad[self.name] = -1.0 * np.dot(Qinv, A - mean) * var_col ** -1

    def prior_kwargs_from(self, kw, parameters, from_mean=False, var=1.0):
        A = getattr(parameters, self.name)
        kw[self.mean] = A.copy() if from_mean else np.zeros_like(A)
        kw[self.var_col] = np.ones(A.shape[1]) * var

    # preconditioner role (matrices.py:632-656, :1094-1125)
    def precondition(self, out, grad, parameters):
        out[self.name] = np.dot(getattr(parameters, self.var_row_name), grad[self.name])

    def precondition_noise(self, out, parameters):
        L = getattr(parameters, "L{0}inv".format(self.var_row_name))
        out[self.name] = np.linalg.solve(L.T, np.random.normal(loc=0, size=np.shape(getattr(parameters, self.name))))

    def correction_term(self, out, parameters):
        out[self.name] = np.zeros_like(getattr(parameters, self.name), dtype=float)


class SquareMatrixVar(_MatrixVar):
    def __init__(self, name="A", dim_name="n", var_row_name=None):
        self._setup(name, var_row_name)
        self.dim_name = dim_name

    def init_param(self, param, **kw):
        if self.name not in kw:
            raise ValueError("{0} not provided".format(self.name))
        n, n2 = np.shape(kw[self.name])
        if n != n2:
            raise ValueError("{0} must be square matrices".format(self.name))
        param.var_dict[self.name] = np.array(kw[self.name]).astype(float)
        param._set_check_dim(**{self.dim_name: n})

    def properties(self):
        return {self.name: _value_prop(self.name), self.dim_name: _dim_prop(self.dim_name)}

    def project(self, param, **kwargs):
        opts = kwargs.get(self.name, {})
        if opts.get("thresh", True):          # matrices.py:465-474
            param.var_dict[self.name] = varp_stability_projection(
                param.var_dict[self.name], eigenvalue_cutoff=opts.get("eigenvalue_cutoff", 0.9999),
                var_name=self.name, logger=logger)
        if opts.get("fixed") is not None:
            param.var_dict[self.name] = opts["fixed"].copy()

    def unflatten(self, out, vector, idx, **dim):
        n = dim[self.dim_name]
        out[self.name] = np.reshape(np.array(vector[idx:idx + n * n], dtype=float), (n, n))
        return idx + n * n

    def init_prior(self, prior, **kw):
        if self.mean not in kw:
            raise ValueError("{0} must be provided".format(self.mean))
        if self.var_col not in kw:
            raise ValueError("{0} must be provided".format(self.var_col))
        n, n2 = np.shape(kw[self.mean])
        if n != n2:
            raise ValueError("{0} must be square".format(self.mean))
        if n != np.size(kw[self.var_col]):
            raise ValueError("prior dimensions don't match")
        prior._set_check_dim(**{self.dim_name: n})
        prior.hyperparams[self.mean] = kw[self.mean]
        prior.hyperparams[self.var_col] = kw[self.var_col]

    def default_prior_kwargs(self, kw, var=100.0, **dims):
        n = dims[self.dim_name]
        kw[self.mean] = np.zeros((n, n))
        kw[self.var_col] = np.ones(n) * var


class RectMatrixVar(_MatrixVar):
    def __init__(self, name="C", dim_names=("m", "n"), var_row_name=None):
        self._setup(name, var_row_name)
        self.dim_names = tuple(dim_names)

    def init_param(self, param, **kw):
        if self.name not in kw:
            raise ValueError("{0} not provided".format(self.name))
        m, n = np.shape(kw[self.name])
        param.var_dict[self.name] = np.array(kw[self.name]).astype(float)
        param._set_check_dim(**{self.dim_names[0]: m, self.dim_names[1]: n})

    def properties(self):
        return {self.name: _value_prop(self.name), self.dim_names[0]: _dim_prop(self.dim_names[0]),
                self.dim_names[1]: _dim_prop(self.dim_names[1])}

    def project(self, param, **kwargs):
        opts = kwargs.get(self.name, {})
        if opts.get("thresh", False):
            param.var_dict[self.name] = varp_stability_projection(
                param.var_dict[self.name], eigenvalue_cutoff=opts.get("eigenvalue_cutoff", 0.9999),
                var_name=self.name, logger=logger)
        if opts.get("fixed") is not None:
            param.var_dict[self.name] = opts["fixed"].copy()
        if opts.get("fixed_eye", False):      # matrices.py:941-945
            k = min(param.dim[self.dim_names[0]], param.dim[self.dim_names[1]])
            A = param.var_dict[self.name]
            A[0:k, 0:k] = np.eye(k)
            param.var_dict[self.name] = A

    def unflatten(self, out, vector, idx, **dim):
        m, n = dim[self.dim_names[0]], dim[self.dim_names[1]]
        out[self.name] = np.reshape(np.array(vector[idx:idx + m * n], dtype=float), (m, n))
        return idx + m * n

    def init_prior(self, prior, **kw):
        if self.mean not in kw:
            raise ValueError("{0} must be provided".format(self.mean))
        if self.var_col not in kw:
            raise ValueError("{0} must be provided".format(self.var_col))
        m, n = np.shape(kw[self.mean])
        if n != np.size(kw[self.var_col]):
            raise ValueError("prior dimensions don't match")
        prior._set_check_dim(**{self.dim_names[0]: m, self.dim_names[1]: n})
        prior.hyperparams[self.mean] = kw[self.mean]
        prior.hyperparams[self.var_col] = kw[self.var_col]

    def default_prior_kwargs(self, kw, var=100.0, **dims):
        m, n = dims[self.dim_names[0]], dims[self.dim_names[1]]
        kw[self.mean] = np.zeros((m, n))
        kw[self.var_col] = np.ones(n) * var


class GARCHVar(object):
    """(log_mu, logit_phi, logit_lambduh) <-> (alpha, beta, gamma)  (garch_var.py:69-91)."""
    names = ("log_mu", "logit_phi", "logit_lambduh")
    hyper = ("scale_mu", "shape_mu", "alpha_phi", "beta_phi", "alpha_lambduh", "beta_lambduh")

    def init_param(self, param, **kw):
        for name in self.names:
            if name not in kw:
                raise ValueError("{0} not provided".format(name))
            param.var_dict[name] = np.atleast_1d(kw[name]).astype(float)

    def properties(self):
        props = {name: _value_prop(name) for name in self.names}
        props["mu"] = property(lambda p: np.exp(p.var_dict["log_mu"]))
        props["phi"] = property(lambda p: expit(p.var_dict["logit_phi"]))
        props["lambduh"] = property(lambda p: expit(p.var_dict["logit_lambduh"]))
        props["alpha"] = property(lambda p: p.mu * (1 - p.phi))
        props["beta"] = property(lambda p: p.phi * p.lambduh)
        props["gamma"] = property(lambda p: p.phi * (1 - p.lambduh))
        return props

    def project(self, param, **kwargs):
        for name in self.names:
            opts = kwargs.get(name, {})
            if opts.get("fixed") is not None:
                param.var_dict[name] = opts["fixed"].copy()

    def flatten(self, var_dict):
        return [var_dict[name].flatten() for name in self.names]

    def unflatten(self, out, vector, idx, **dim):
        for name in self.names:
            out[name] = np.reshape(np.array(vector[idx:idx + 1], dtype=float), (1,))
            idx += 1
        return idx

    def init_prior(self, prior, **kw):
        for name in self.hyper:
            if name not in kw:
                raise ValueError("{0} must be provided".format(name))
            prior.hyperparams[name] = kw[name]

    def sample_prior(self, prior, var_dict):
        h = prior.hyperparams
        var_dict["log_mu"] = np.log(scipy.stats.invgamma(a=h["shape_mu"], scale=h["scale_mu"]).rvs())
        var_dict["logit_phi"] = logit(scipy.stats.beta(a=h["alpha_phi"], b=h["beta_phi"]).rvs())
        var_dict["logit_lambduh"] = logit(scipy.stats.beta(a=h["alpha_lambduh"], b=h["beta_lambduh"]).rvs())

    def logprior(self, prior, parameters):
        h = prior.hyperparams
        out = scipy.stats.invgamma(a=h["shape_mu"], scale=h["scale_mu"]).logpdf(parameters.mu)
        out = out + scipy.stats.beta(a=h["alpha_phi"], b=h["beta_phi"]).logpdf((1 + parameters.phi) / 2.0)
        out = out + scipy.stats.beta(a=h["alpha_lambduh"], b=h["beta_lambduh"]).logpdf((1 + parameters.lambduh) / 2.0)
        return float(np.ravel(out)[0])

    def grad_logprior(self, prior, grad, parameters):
        h = prior.hyperparams                 # garch_var.py:150-163
        grad["log_mu"] = -h["shape_mu"] - 1 + h["scale_mu"] / parameters.mu
        phi, lam = parameters.phi, parameters.lambduh
        grad["logit_phi"] = ((h["alpha_phi"] - 1) / (1 + phi) - (h["beta_phi"] - 1) / (1 - phi)) * phi * (1 - phi)
        grad["logit_lambduh"] = ((h["alpha_lambduh"] - 1) / (1 + lam) - (h["beta_lambduh"] - 1) / (1 - lam)) * lam * (1 - lam)

    def default_prior_kwargs(self, kw, var=100.0, **dims):
        var = min(var, 1)                     # garch_var.py:178-188
        kw["scale_mu"] = var + 2
        kw["shape_mu"] = kw["scale_mu"] + 1
        kw["alpha_phi"] = 1 + 19 * var ** -1
        kw["beta_phi"] = kw["alpha_phi"] / 9
        kw["alpha_lambduh"] = 1 + 19 * var ** -1
        kw["beta_lambduh"] = kw["alpha_lambduh"] / 9

    def prior_kwargs_from(self, kw, parameters, from_mean=False, var=1.0):
        self.default_prior_kwargs(kw, var=var)
