"""Oracle functional kernels (hand-derived, closed shell) vs. torch.autograd of the spin-resolved expressions."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from functional_reference import BY_ID, closed_shell


def _points(seed=3, n=400):
    rng = np.random.default_rng(seed)
    rho = 10.0 ** rng.uniform(-9, 2.3, size=n)
    s = 10.0 ** rng.uniform(-3, 1.0, size=n)          # reduced gradient
    sigma = (s * 2.0 * (3 * np.pi ** 2) ** (1 / 3) * rho ** (4 / 3)) ** 2
    return rho, sigma


@pytest.mark.parametrize("fid", sorted(BY_ID))
def test_oracle_kernel_matches_autograd(fid):
    rho, sigma = _points()
    F, vr, vs = closed_shell(fid, rho, sigma)
    got = np.array([orc.basic_functional(fid, r, s) for r, s in zip(rho, sigma)])
    for k, (ref, name) in enumerate(zip((F, vr, vs), ("F", "vrho", "vsigma"))):
        scale = np.maximum(np.abs(ref), 1e-300)
        err = np.abs(got[:, k] - ref) / scale
        # contributions enter E and V weighted by rho-sized factors; 1e-11 relative is far inside 1e-9 Eh
        assert np.all((err < 2e-11) | (np.abs(got[:, k] - ref) < 1e-18)), (name, fid, float(err.max()))


def test_functional_on_grid_thresholds():
    """Block skip (sum|rho| < n*1e-12) and tiny-density zeroing (rho < 1e-14), XCFun.cpp:135-140."""
    f = orc.Functional([135, 197], [1.0, 1.0])
    n = 300
    rho = np.full(n, 1e-3)
    rho[:128] = 5e-13          # first block skipped as a whole
    rho[130] = 5e-15           # single tiny point in an evaluated block
    g = [np.full(n, 1e-4) for _ in range(3)]
    w = np.ones(n)
    e, out = orc.functional_on_grid(f, w, rho, *g)
    assert np.all(out[0][:128] == 0) and np.all(out[1][:128] == 0) and np.all(out[2][:128] == 0)
    assert out[0][130] == 0 and out[1][130] == 0
    assert out[0][129] != 0 and out[1][200] != 0 and out[2][200] != 0
    assert abs(e - out[0].sum()) < 1e-15


def test_composite_is_linear_combination():
    ids, mix = [2, 81, 184, 45], [0.80, 0.72, 0.81, 0.19]   # B3LYP, CompositeFunctionals.cpp:272-276
    rho, sigma = _points(seed=5, n=256)
    gx = np.sqrt(sigma)
    z = np.zeros_like(gx)
    e, out = orc.functional_on_grid(orc.Functional(ids, mix), np.ones_like(rho), rho, gx, z, z)
    F = sum(m * closed_shell(i, rho, sigma)[0] for i, m in zip(ids, mix))
    vs = sum(m * closed_shell(i, rho, sigma)[2] for i, m in zip(ids, mix))
    assert np.allclose(out[0], F, rtol=1e-11, atol=1e-18)
    assert np.allclose(out[2], 2 * vs * gx, rtol=1e-10, atol=1e-18)
