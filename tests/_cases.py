"""Helpers to read tests/golden/*.npz fixtures."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return _cache[name]


def case_names(section, file="ref_cases.npz"):
    z = load(file)
    names = sorted({"/".join(k.split("/")[:2]) for k in z.files if k.startswith(section + "/")})
    return names


def case(name, file="ref_cases.npz"):
    z = load(file)
    pre = name + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


THETA_KEYS = {
    "svm": ["A", "LQinv", "Qinv", "LRinv", "Rinv"],
    "lgssm": ["A", "LQinv", "Qinv", "C", "LRinv", "Rinv"],
    "garch": ["alpha", "beta", "gamma", "mu", "phi", "lambduh", "LRinv", "Rinv", "R"],
}


def theta_dict(model, vec):
    return {k: float(v) for k, v in zip(THETA_KEYS[model], np.ravel(vec))}


def parse_kernel_case(name):
    """'k/<model>_<kernel>_<pf...>_<N>_<opts>' -> (model, kernel, pf)."""
    body = name.split("/", 1)[1]
    model, kernel, rest = body.split("_", 2)
    for pf in ("poyiadjis_N2", "poyiadjis_N", "nemeth", "paris", "filter"):
        if rest.startswith(pf + "_"):
            return model, kernel, pf
    raise ValueError(name)


def case_opts(c):
    opts = {}
    for k, v in c.items():
        if k.startswith("opt_"):
            v = v.item()
            opts[k[4:]] = v
    return opts
