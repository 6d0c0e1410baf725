"""The CPU oracle against the reference's own known-answer vectors (SURVEY.md section 8c i-iii).

Golden files were extracted from the reference's unit tests by tests/golden/make_golden_from_reference.py.
"""
import math

import numpy as np

from conftest import grid_arrays, load_golden
from oracle import pyoracle as orc


def test_renormalisation_constants():
    """Appendix B: test values are divided by sqrt(1.1930904 / 1.1826193 / 1.2623580)
    (BasisFunctionOnGridController_test.cpp:84-123) - the unit-normalisation N^2 of the three shells."""
    from serenity_b200.inputs.basis import renormalise
    shells = [(0, [1.0, 2.0], [1.0, 0.1]), (1, [1.0, 2.0], [1.0, 0.1]), (2, [0.5, 1.2], [0.8, 0.4])]
    for (l, exps, coefs), n2 in zip(shells, [1.1930904, 1.1826193, 1.2623580]):
        a = np.asarray(exps)
        df = {0: 1.0, 1: 1.0, 2: 3.0}[l]
        prim = np.asarray(coefs) * np.sqrt(2.0 ** l * (2 * a) ** (l + 1.5) / (math.pi ** 1.5 * df))
        ratio = prim / renormalise(l, exps, coefs)
        assert np.allclose(ratio ** 2, n2, rtol=0, atol=5e-8)


def test_basis_function_values_and_derivatives(fixtures, small_mixed):
    """BasisFunctionOnGridController_test.cpp:43-483: phi, grad phi, hess phi on TINY x SMALL_MIXED @1e-8."""
    gold = load_golden("basis_functions_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    basis, grid = orc.Basis(small_mixed), orc.Grid(xyz, w, gold["block_size"])
    arrs, neg, _ = orc.basis_block(basis, grid, gold["radial_threshold"], 2, 0)
    names = ["values", "dx", "dy", "dz", "hxx", "hxy", "hxz", "hyy", "hyz", "hzz"]
    n_checked = 0
    for name, arr in zip(names, arrs):
        for p, mu, ref in gold["entries"][name]:
            assert abs(arr[p, mu] - ref) < gold["tolerance"], (name, p, mu, arr[p, mu], ref)
            n_checked += 1
    assert n_checked == 400


def test_density_gradient_hessian(fixtures, small_mixed):
    """DensityOnGridCalculator_test.cpp:43-254: rho, grad rho, hess rho @5e-8 (block size 3, threshold 0)."""
    gold = load_golden("density_ref.json")
    xyz, w = grid_arrays(fixtures, "TINY")
    basis, grid = orc.Basis(small_mixed), orc.Grid(xyz, w, gold["block_size"])
    rho, g, h, nonneg = orc.density_on_grid(basis, grid, gold["radial_threshold"], np.asarray(gold["P"]), deriv=2)
    exp = gold["expected"]
    tol = gold["tolerance"]
    assert np.allclose(rho, exp["rho"], rtol=0, atol=tol)
    for k, c in enumerate("xyz"):
        assert np.allclose(g[k], exp["d" + c], rtol=0, atol=tol), c
    for k, c in enumerate(["xx", "xy", "xz", "yy", "yz", "zz"]):
        assert np.allclose(h[k], exp["h" + c], rtol=0, atol=tol), c
    assert nonneg.tolist() == [1, 1]


def test_scalar_and_gradient_operator_to_matrix(fixtures, small_mixed):
    """ScalarOperatorToMatrixAdder_test.cpp:41-148: 55 matrix elements @1e-8."""
    gold = load_golden("scatter_ref.json")
    xyz, w = grid_arrays(fixtures, "VERY_SMALL")
    basis, grid = orc.Basis(small_mixed), orc.Grid(xyz, w, gold["block_size"])
    pot = gold["potential"]
    V = orc.scalar_to_matrix(basis, grid, gold["radial_threshold"], gold["block_ave_threshold"],
                             np.asarray(pot["pot"]), np.asarray(pot["gx"]), np.asarray(pot["gy"]), np.asarray(pot["gz"]))
    assert len(gold["entries"]) == 55
    for i, j, ref in gold["entries"]:
        assert abs(V[i, j] - ref) < gold["tolerance"], (i, j, V[i, j], ref)
        assert V[i, j] == V[j, i]


def test_harmonics_against_reference_formulas():
    """Generated solid harmonics vs. the reference's hard-coded ones (BasisFunctionOnGridController.cpp:441-1066),
    value / gradient / Hessian for l = 0..6, through a single-primitive spherical shell."""
    gold = load_golden("harmonics_ref.json")
    from serenity_b200.inputs.basis import shell_table_from_list
    pts = np.asarray(gold["points"])
    a = 0.3
    for l in range(7):
        tab = shell_table_from_list([{"l": l, "pure": True, "exps": [a], "coefs": [1.0], "centre": [0.0, 0.0, 0.0]}])
        c = tab.coeff[0]
        grid = orc.Grid(pts, np.ones(len(pts)), 128)
        arrs, neg, _ = orc.basis_block(orc.Basis(tab), grid, 1e-30, 2, 0)
        for ip, p in enumerate(pts):
            r2 = float(p @ p)
            R = c * np.exp(-a * r2)
            dR = -2 * a * R
            ddR = 4 * a * a * R
            ref = dict(zip(gold["names"], [np.asarray(v) for v in gold["values"][str(l)][ip]]))
            Y = ref["Y"]
            x, y, z = p
            exp_val = R * Y
            exp_d = [R * ref["dYdx"] + dR * x * Y, R * ref["dYdy"] + dR * y * Y, R * ref["dYdz"] + dR * z * Y]
            exp_h = [R * ref["d2Ydxdx"] + 2 * dR * x * ref["dYdx"] + ddR * x * x * Y + dR * Y,
                     R * ref["d2Ydxdy"] + dR * x * ref["dYdy"] + dR * y * ref["dYdx"] + ddR * x * y * Y,
                     R * ref["d2Ydxdz"] + dR * x * ref["dYdz"] + dR * z * ref["dYdx"] + ddR * x * z * Y,
                     R * ref["d2Ydydy"] + 2 * dR * y * ref["dYdy"] + ddR * y * y * Y + dR * Y,
                     R * ref["d2Ydydz"] + dR * y * ref["dYdz"] + dR * z * ref["dYdy"] + ddR * y * z * Y,
                     R * ref["d2Ydzdz"] + 2 * dR * z * ref["dYdz"] + ddR * z * z * Y + dR * Y]
            for got, exp in zip([arrs[0]] + arrs[1:4] + arrs[4:], [exp_val] + exp_d + exp_h):
                assert np.allclose(got[ip], exp, rtol=1e-12, atol=1e-13), (l, ip)
