"""pytest configuration: marker registration and import paths.

`-m "not gpu"` runs the CPU suite (oracle vs golden vectors, host logic, C-ABI symbol check);
`-m gpu` runs the CUDA parity tests through the C-ABI on a B200.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a)")
