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


# ------------------------------------------------------------------------------------------- UNRESTRICTED
def _spin_points(seed=7, n=300):
    rng = np.random.default_rng(seed)
    rho = 10.0 ** rng.uniform(-8, 2.0, size=n)
    zeta = rng.uniform(-0.95, 0.95, size=n)
    ra, rb = 0.5 * rho * (1 + zeta), 0.5 * rho * (1 - zeta)
    s = 10.0 ** rng.uniform(-3, 0.8, size=n)
    gnorm = s * 2.0 * (3 * np.pi ** 2) ** (1 / 3) * rho ** (4 / 3)
    ua, ub = rng.normal(size=(3, n)), rng.normal(size=(3, n))
    ga = 0.5 * gnorm * (1 + zeta) * ua / np.linalg.norm(ua, axis=0)
    gb = 0.5 * gnorm * (1 - zeta) * ub / np.linalg.norm(ub, axis=0)
    return ra, rb, ga, gb


@pytest.mark.parametrize("fid", sorted(BY_ID))
def test_oracle_spin_polarised_kernel_matches_autograd(fid):
    from functional_reference import spin_resolved
    ra, rb, ga, gb = _spin_points()
    gaa, gab, gbb = (ga * ga).sum(0), (ga * gb).sum(0), (gb * gb).sum(0)
    F, d = spin_resolved(fid, ra, rb, gaa, gab, gbb)
    for p in range(len(ra)):
        Fo, do = orc.basic_functional_u(fid, ra[p], rb[p], gaa[p], gab[p], gbb[p])
        assert abs(Fo - F[p]) <= 2e-11 * abs(F[p]) + 1e-18
        for k in range(5):
            # a derivative that is a small difference of large terms is judged on the scale of its siblings
            scale = max(abs(d[k, p]), 1e-3 * np.abs(d[:2, p]).max() if k < 2 else 1e-3 * np.abs(d[2:, p]).max())
            assert abs(do[k] - d[k, p]) <= 5e-11 * scale + 1e-18, (fid, p, k, do[k], d[k, p])


def test_unrestricted_functional_on_grid_reduces_to_restricted():
    """rho_a = rho_b = rho/2: epuv and E equal the restricted call; dF/drho_a = dF/drho, dF/dgrad_a = dF/dgrad."""
    ids, mix = [135, 197], [1.0, 1.0]
    rng = np.random.default_rng(9)
    n = 300
    rho = 10.0 ** rng.uniform(-6, 1, size=n)
    rho[:128] = 3e-13                                  # skipped block
    g = rng.normal(size=(3, n)) * rho ** (4 / 3)
    w = rng.uniform(0.5, 1.5, size=n)
    f = orc.Functional(ids, mix)
    e_r, out = orc.functional_on_grid(f, w, rho, *[np.ascontiguousarray(x) for x in g])
    e_u, ep, vr, vg = orc.functional_on_grid_u(f, w, np.stack([rho / 2, rho / 2]), np.stack([g / 2, g / 2]))
    assert abs(e_r - e_u) < 1e-12 * abs(e_r)
    assert np.allclose(ep, out[0], rtol=1e-11, atol=1e-18) and np.all(ep[:128] == 0)
    for s in range(2):
        assert np.allclose(vr[s], out[1], rtol=1e-10, atol=1e-16)
        for c in range(3):
            assert np.allclose(vg[s, c], out[2 + c], rtol=1e-9, atol=1e-16)
