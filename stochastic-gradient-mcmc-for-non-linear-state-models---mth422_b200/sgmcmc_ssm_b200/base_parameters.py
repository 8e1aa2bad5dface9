"""Parameters / Prior / Preconditioner containers.

API mirror of sgmcmc_ssm/base_parameters.py:12-297 (same public methods and attribute names:
`var_dict`, `dim`, `as_dict`, `as_vector`, `from_vector`, `copy`, `project_parameters`,
`sample_prior`, `logprior`, `grad_logprior`, `generate_default_prior`, `generate_prior`,
`precondition`, `precondition_noise`, `correction_term`).  Unlike the reference's three parallel
helper hierarchies, each *variable* (variables.py) carries its parameter, prior and preconditioner
behaviour in one object; containers just iterate over their variable list.
"""
from copy import deepcopy

import numpy as np


class _Dims(object):
    def _set_check_dim(self, **kwargs):
        for k, v in kwargs.items():
            if k in self.dim and self.dim[k] != v:
                raise ValueError("{0} does not match existing dims {1} != {2}".format(k, v, self.dim[k]))
            self.dim[k] = v


class BaseParameters(_Dims):
    _variables = []          # ordered: defines var_dict / vector order

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        for var in cls._variables:
            for name, prop in var.properties().items():
                setattr(cls, name, prop)

    def __init__(self, **kwargs):
        self.dim = {}
        self.var_dict = {}
        for var in self._variables:
            var.init_param(self, **kwargs)

    def as_dict(self, copy=True):
        return self.var_dict.copy() if copy else self.var_dict

    def as_vector(self):
        return self.from_dict_to_vector(self.var_dict, **self.dim)

    def from_vector(self, vector):
        self.var_dict.update(self.from_vector_to_dict(vector, **self.dim))

    # `parameters.vector` is what metric_functions.average_input_decorator reads and assigns (metric_functions.py:239-261
    # of the reference, whose containers lack the attribute: the decorator raises AttributeError there)
    vector = property(lambda self: self.as_vector(), lambda self, value: self.from_vector(np.asarray(value, dtype=float)))

    @classmethod
    def from_dict_to_vector(cls, var_dict, **dim):
        return np.concatenate([np.ravel(v) for var in cls._variables for v in var.flatten(var_dict)])

    @classmethod
    def from_vector_to_dict(cls, vector, **dim):
        out, idx = {}, 0
        for var in cls._variables:
            idx = var.unflatten(out, vector, idx, **dim)
        return out

    def __iadd__(self, other):
        if not isinstance(other, dict):
            raise TypeError("Addition only defined for dict not {0}".format(type(other)))
        for key in self.var_dict:
            self.var_dict[key] += other[key]
        return self

    def __add__(self, other):
        out = self.copy()
        out += other
        return out

    __radd__ = __add__

    def copy(self):
        return type(self)(**deepcopy(self.var_dict))

    def project_parameters(self, **kwargs):
        for var in self._variables:
            var.project(self, **kwargs)
        return self


class BasePrior(_Dims):
    _Parameters = BaseParameters
    _variables = []          # ordered: defines sampling order (a matrix needs its row covariance first)

    def __init__(self, **kwargs):
        self.dim = {}
        self.hyperparams = {}
        for var in self._variables:
            var.init_prior(self, **kwargs)

    def sample_prior(self, **kwargs):
        var_dict = {}
        for var in self._variables:
            var.sample_prior(self, var_dict)
        return self._Parameters(**var_dict)

    def logprior(self, parameters, **kwargs):
        return float(sum(var.logprior(self, parameters) for var in self._variables))

    def grad_logprior(self, parameters, **kwargs):
        grad = {}
        for var in self._variables:
            var.grad_logprior(self, grad, parameters)
        return grad

    @classmethod
    def generate_prior(cls, parameters, from_mean=False, var=1.0):
        kw = {}
        for v in cls._variables:
            v.prior_kwargs_from(kw, parameters, from_mean=from_mean, var=var)
        return cls(**kw)

    @classmethod
    def generate_default_prior(cls, var=100.0, **dims):
        kw = {}
        for v in cls._variables:
            v.default_prior_kwargs(kw, var=var, **dims)
        return cls(**kw)


class BasePreconditioner(object):
    _variables = []

    def __init__(self, **kwargs):
        pass

    def precondition(self, grad, parameters, scale=1.0, **kwargs):
        out = {}
        for var in self._variables:
            var.precondition(out, grad, parameters)
        for k in out:
            out[k] *= scale
        return out

    def precondition_noise(self, parameters, scale=1.0):
        out = {}
        for var in self._variables:
            var.precondition_noise(out, parameters)
        for k in out:
            out[k] *= scale ** 0.5
        return out

    def correction_term(self, parameters, scale=1.0):
        out = {}
        for var in self._variables:
            var.correction_term(out, parameters)
        for k in out:
            out[k] = out[k] * scale
        return out
