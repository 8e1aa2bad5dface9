"""Device check of the table-driven logarithm of the f64 variate transforms (csrc/fastlog.cuh) through the C-ABI self-test hook."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_fast_log_matches_the_host_logarithm():
    import torch
    from sgmcmc_ssm_b200 import _native
    lib = _native.load()
    rs = np.random.RandomState(3)
    x = np.concatenate([rs.uniform(0, 1, 1 << 20), 1 - rs.uniform(0, 1, 4096) * 1e-9, rs.uniform(0, 1, 4096) * 1e-15,
                        [1 - 2.0 ** -53, 2.0 ** -52, 0.5, 0.25, 2.0 ** -0.5, (0.5 + 2.0 ** -30) * 2.0 ** -52]])
    x = x[(x > 0) & (x < 1)]
    xd = torch.from_numpy(x).cuda()
    yd = torch.empty_like(xd)
    _native.check(lib.sgm_selftest_log(xd.data_ptr(), yd.data_ptr(), xd.numel(), None))
    torch.cuda.synchronize()
    y = yd.cpu().numpy()
    ref = np.log(x)
    rel = np.abs(y - ref) / np.abs(ref)
    assert rel.max() <= 4 * 2.0 ** -53, (rel.max(), x[rel.argmax()])
    # spacings must be non-negative: -log(u) >= 0 for every u < 1
    assert (y <= 0).all()
