#!/usr/bin/env python3
"""Generate tests/golden/ksd_cases.npz: the unmodified reference's IMQ_KSD (trace_metric_functions.py:20-81) on
seeded synthetic traces (incl. the blocked path, num_points > max_block_size).  Build-container only."""
import os, sys, warnings
import numpy as np
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")
from sgmcmc_ssm.trace_metric_functions import IMQ_KSD  # noqa: E402
rs = np.random.RandomState(0)
out = {}
for tag, (K, d, c, beta) in {"a": (37, 1, 1, 0.5), "b": (1500, 3, 1, 0.5), "c": (2300, 4, 0.7, 0.3), "d": (256, 8, 2.0, 0.9)}.items():
    x = rs.normal(size=(K, d)); g = -x + 0.3 * rs.normal(size=(K, d))
    for k, v in dict(x=x, g=g, c=c, beta=beta, out=IMQ_KSD(x, g, c=c, beta=beta)).items():
        out["ksd/%s/%s" % (tag, k)] = np.asarray(v)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ksd_cases.npz"), **out)
