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
    _native.check(lib.sgm_selftest_math(0, xd.data_ptr(), yd.data_ptr(), xd.numel(), None))
    torch.cuda.synchronize()
    y = yd.cpu().numpy()
    ref = np.log(x)
    rel = np.abs(y - ref) / np.abs(ref)
    assert rel.max() <= 4 * 2.0 ** -53, (rel.max(), x[rel.argmax()])
    # spacings must be non-negative: -log(u) >= 0 for every u < 1
    assert (y <= 0).all()


def test_device_fast_sincos_matches_the_host_functions():
    import torch
    from sgmcmc_ssm_b200 import _native
    lib = _native.load()
    rs = np.random.RandomState(5)
    bits = rs.randint(0, 2 ** 63, size=1 << 18, dtype=np.int64).astype(np.uint64) * np.uint64(2) + rs.randint(0, 2, size=1 << 18).astype(np.uint64)
    bits[:6] = np.array([0, 2 ** 64 - 1, 1 << 62, 1 << 63, 3 << 62, (1 << 56) - 1], dtype=np.uint64)
    xd = torch.from_numpy(bits.view(np.int64).copy()).cuda()
    yd = torch.empty(2 * bits.size, dtype=torch.float64, device="cuda")
    _native.check(lib.sgm_selftest_math(1, xd.data_ptr(), yd.data_ptr(), bits.size, None))
    torch.cuda.synchronize()
    y = yd.cpu().numpy().reshape(-1, 2)
    # v = (k + f) / 256 from the top 60 bits; reference in x87 extended precision (64-bit mantissa holds v exactly)
    assert np.finfo(np.longdouble).eps < 2.0 ** -60
    v = (bits >> np.uint64(4)).astype(np.longdouble) / np.longdouble(2.0) ** 60
    ang = 8 * np.arctan(np.longdouble(1.0)) * v
    ref_s, ref_c = np.sin(ang).astype(np.float64), np.cos(ang).astype(np.float64)
    assert np.abs(y[:, 0] - ref_s).max() <= 1e-15 and np.abs(y[:, 1] - ref_c).max() <= 1e-15
    assert np.abs(y[:, 0] ** 2 + y[:, 1] ** 2 - 1.0).max() <= 2e-15
