"""Every kernel family x model x resampling scheme on small ragged batches (scripts/all_kernels_smoke.py)."""
import os
import runpy

import pytest

pytestmark = pytest.mark.gpu


def test_all_kernel_families_run_and_stay_finite():
    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "all_kernels_smoke.py"))
