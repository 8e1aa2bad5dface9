"""The f64 variate logarithm (csrc/fastlog.cuh): its generated table is current, and the algorithm -- restated here with exact
rational arithmetic for every fused multiply-add -- is within 4 ulp of math.log on (0, 1), with full relative accuracy at 1."""
import math
import os
import sys
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import make_log_table as mlt  # noqa: E402

PKG = os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200")


def host_fast_log(x, tab):
    m, e = math.frexp(x)
    m *= 2.0
    e -= 1
    idx = int((np.float64(m).view(np.int64) >> 45) & 127)
    a, b = tab[idx]
    fma = lambda p, q, r: float(Fraction(p) * Fraction(q) + Fraction(r))
    r = fma(m, a, -1.0)
    p = 1.0 / 7.0
    for c in (-1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5, 1.0):
        p = fma(p, r, c)
    ee = e + (1 if idx >= mlt.FOLD else 0)
    return fma(p, r, fma(float(ee), 0.6931471805599453, b))


def test_committed_table_is_what_the_generator_writes():
    with open(os.path.join(PKG, "csrc", "log_table.cuh")) as f:
        assert f.read() == mlt.render()


def test_algorithm_is_within_4_ulp_on_the_unit_interval():
    tab = [mlt.entry(i) for i in range(128)]
    rs = np.random.RandomState(7)
    xs = np.concatenate([rs.uniform(0, 1, 3000), 1 - rs.uniform(0, 1, 300) * 1e-6, rs.uniform(0, 1, 300) * 1e-12,
                         1 - rs.uniform(0, 1, 300) * 1e-2, [1 - 2.0 ** -53, 2.0 ** -52, 0.5, 0.25, 2.0 ** -0.5]])
    worst = 0.0
    for x in xs:
        if 0.0 < x < 1.0:
            worst = max(worst, abs(host_fast_log(float(x), tab) - math.log(x)) / abs(math.log(x)))
    assert worst <= 4 * 2.0 ** -53, worst
