"""Evaluation-tick callers of the PF path: factories returning `f(sampler) -> dict | list[dict]` records
(`variable`, `metric`, `value`), the protocol of the reference's evaluators (reference
`sgmcmc_ssm/metric_functions.py`).  Covered: the generators the LGSSM / SVM / GARCH drivers use around
`kind='pf'` -- the noisy log-joint / log-likelihood pair (`:362-381`), the k-step-ahead predictive
log-likelihood (`:383-417`), parameter samples and parameter-vs-target metrics (`:8-177`, `:205-237`),
`metric_function_from_sampler` (`:179-203`) and the running-average decorator (`:239-261`).  The label-permutation
variants (`:263-360`) and the z / x comparisons (`:419-`) belong to the discrete-state models, which are out of
scope."""
import numpy as np

_METRICS = {
    "mse": lambda r, e: np.mean((r - e) ** 2),
    "logmse": lambda r, e: np.log10(np.mean((r - e) ** 2)),
    "rmse": lambda r, e: np.sqrt(np.mean((r - e) ** 2)),
    "mae": lambda r, e: np.mean(np.abs(r - e)),
}


def construct_metric_function(metric_name):
    """`(result, expected) -> float` for 'mse' | 'logmse' | 'rmse' | 'mae' (metric_functions.py:205-237)."""
    try:
        return _METRICS[metric_name]
    except KeyError:
        raise ValueError("Unrecognized metric name = %s" % metric_name)


def _names(parameter_names, return_variable_names):
    if return_variable_names is None:
        return list(parameter_names)
    if len(return_variable_names) != len(parameter_names):
        raise ValueError("parameter and return names must be equal length")
    return [p if r is None else r for p, r in zip(parameter_names, return_variable_names)]


def sample_function_parameter(parameter_name, return_variable_name=None):
    """Record holding a copy of `sampler.parameters.<parameter_name>` (metric_functions.py:8-30)."""
    variable = parameter_name if return_variable_name is None else return_variable_name

    def custom_sample_function(sampler):
        return {"variable": variable, "value": np.copy(getattr(sampler.parameters, parameter_name))}
    return custom_sample_function


def sample_function_parameters(parameter_names, return_variable_names=None, decorator=None):
    """List of parameter records (metric_functions.py:32-66)."""
    fns = [sample_function_parameter(p, v) for p, v in zip(parameter_names, _names(parameter_names, return_variable_names))]

    def custom_sample_function(sampler):
        return [f(sampler) for f in fns]
    return custom_sample_function if decorator is None else decorator(custom_sample_function)


def metric_function_parameter(parameter_name, target_value, metric_name, return_variable_name=None):
    """Distance of the current parameter to a target (metric_functions.py:68-100)."""
    dist = construct_metric_function(metric_name)
    variable = parameter_name if return_variable_name is None else return_variable_name

    def custom_metric_function(sampler):
        return {"variable": variable, "metric": metric_name,
                "value": dist(getattr(sampler.parameters, parameter_name), target_value)}
    return custom_metric_function


def metric_function_parameters(parameter_names, target_values, metric_names, return_variable_names=None,
                               decorator=None, criteria=None, double_permutation_flag=False):
    """List of parameter-vs-target metrics (metric_functions.py:102-177).  `criteria` selects the label-permutation
    search of the discrete-state models and is not available here."""
    if len(target_values) != len(parameter_names) or len(metric_names) != len(parameter_names):
        raise ValueError("input args not equal length")
    if criteria is not None:
        raise NotImplementedError("label-permutation metrics belong to the discrete-state models (out of scope)")
    fns = [metric_function_parameter(p, t, m, v) for p, t, m, v in
           zip(parameter_names, target_values, metric_names, _names(parameter_names, return_variable_names))]

    def custom_metric_function(sampler):
        return [f(sampler) for f in fns]
    return custom_metric_function if decorator is None else decorator(custom_metric_function)


def metric_function_from_sampler(sampler_func_name, metric_name=None, return_variable_name="sampler",
                                 **sampler_func_kwargs):
    """Metric = the value of a sampler method, e.g. `noisy_loglikelihood(kind='pf', N=5000)`
    (metric_functions.py:179-203)."""
    metric_name = sampler_func_name if metric_name is None else metric_name

    def custom_metric_function(sampler):
        fn = getattr(sampler, sampler_func_name, None)
        if fn is None:
            raise ValueError("sampler_func_name `{}` is not in sampler".format(sampler_func_name))
        return {"variable": return_variable_name, "metric": metric_name, "value": fn(**sampler_func_kwargs)}
    return custom_metric_function


def average_input_decorator(sampler_function):
    """Evaluate `sampler_function` at the running mean of the parameter vectors seen so far; the sampler's own
    parameters are restored afterwards (metric_functions.py:239-261)."""
    def average_function(sampler):
        average_function.num_calls += 1
        average_function.sum_vector = average_function.sum_vector + sampler.parameters.vector
        current = sampler.parameters.vector
        sampler.parameters.vector = average_function.sum_vector / average_function.num_calls
        try:
            output = sampler_function(sampler)
        finally:
            sampler.parameters.vector = current
        for rec in ([output] if isinstance(output, dict) else output):
            rec["variable"] = "avg_" + rec["variable"]
        return output
    average_function.num_calls = 0
    average_function.sum_vector = 0.0
    return average_function


def noisy_logjoint_loglike_metric(metric_name_prefix="", **kwargs):
    """Two records per tick, `<prefix>noisy_logjoint` and `<prefix>noisy_loglikelihood`, from ONE
    `sampler.noisy_logjoint(return_loglike=True, **kwargs)` -- with `kind='pf'` one batched filter launch
    (metric_functions.py:362-381; the drivers pass `kind='pf', N=5000`, e.g. `models/svm/driver.py:602`)."""
    names = [metric_name_prefix + "noisy_logjoint", metric_name_prefix + "noisy_loglikelihood"]

    def custom_metric_func(sampler):
        res = sampler.noisy_logjoint(return_loglike=True, **kwargs)
        return [dict(variable="sampler", metric=names[0], value=res["logjoint"]),
                dict(variable="sampler", metric=names[1], value=res["loglikelihood"])]
    return custom_metric_func


def noisy_predictive_logjoint_loglike_metric(num_steps_ahead, kind="marginal", metric_name_prefix="", **kwargs):
    """`<prefix><k>_pred_loglikelihood`, k = 0..num_steps_ahead, for `kind='pf'` (one record per horizon); only the
    last horizon otherwise (metric_functions.py:383-417).

    Like the reference, the horizon is forwarded as `lag=` (`:394-397`), which the pf branch of
    `predictive_loglikelihood` does not read: the filter runs its default 10 horizons and the first
    `num_steps_ahead + 1` are reported (so `num_steps_ahead` > 10 raises IndexError, as there).  Kept as is because
    the reference's all-horizon log-sum (`pf.py:73-76`) makes every reported value depend on the number of horizons
    computed.  For per-horizon values call `sampler.predictive_loglikelihood(kind='pf', num_steps_ahead=k,
    per_horizon=True)` directly (or through `metric_function_from_sampler`)."""
    names = ["{0}{1}_pred_loglikelihood".format(metric_name_prefix, k) for k in range(num_steps_ahead + 1)]

    def custom_metric_func(sampler):
        res = sampler.predictive_loglikelihood(lag=num_steps_ahead, kind=kind, **kwargs)
        if kind == "pf":
            return [dict(variable="sampler", metric=names[k], value=res[k]) for k in range(num_steps_ahead + 1)]
        return [dict(variable="sampler", metric=names[-1], value=res)]
    return custom_metric_func
