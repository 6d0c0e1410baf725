"""Grid construction: the partition-weight step of GridFactory::produce
(src/grid/construction/GridFactory.cpp:139-266).

Known answers: the reference's own GridFactory_test.cpp:64-190 integrates the indicator function of a 7-bohr sphere
over the C60 grid for eight (flavour, smoothing, accuracy) combinations and pins the result to narrow windows
(e.g. BECKE/3/acc 4: 1 + 8e-4 < I/V < 1 + 9e-4).  tests/golden/grid_kats.json holds the geometry and the windows
(tests/golden/make_grid_kats.py).  The host restatement (serenity_b200/inputs: atom grids, gridweights.c) is checked
against all eight on CPU; the device kernel (sxc_partition_weights) is compared with the host restatement point by
point and replayed against the same windows.
"""
import math

import numpy as np
import pytest

from conftest import load_golden


def _kats():
    d = load_golden("grid_kats.json")
    return np.asarray(d["c60_bohr"]), d["sphere_radius_bohr"], d["cases"]


def _sphere_ratio(xyz, w, radius):
    inside = np.linalg.norm(xyz, axis=1) <= radius
    return w[inside].sum() / (4.0 / 3.0 * math.pi * radius ** 3)


def _check_window(case, ratio):
    assert abs(ratio - 1.0) < case["near"], (case["test"], ratio)
    if case["greater_than_1_plus"] is not None:
        assert ratio > 1.0 + case["greater_than_1_plus"], (case["test"], ratio)
    if case["less_than_1_minus"] is not None:
        assert ratio < 1.0 - case["less_than_1_minus"], (case["test"], ratio)


def test_host_restatement_meets_the_reference_grid_kats():
    from serenity_b200.inputs.grid import molecular_grid
    c60, radius, cases = _kats()
    assert len(cases) == 8
    for c in cases:
        xyz, w = molecular_grid(["C"] * 60, c60, c["accuracy"], flavour=c["flavour"], radial=c["radial"],
                                weight_threshold=c["weight_threshold"], sort=False, smoothing=c["smoothing"])
        _check_window(c, _sphere_ratio(xyz, w, radius))


def test_single_atom_keeps_atomic_weights_and_partition_is_a_partition_of_unity():
    from serenity_b200.inputs.grid import host_partition_weights, reference_atom_grids
    xyz, w0, parent = reference_atom_grids(["O"], np.zeros((1, 3)), 2)
    assert np.array_equal(host_partition_weights("SSF", [8], np.zeros((1, 3)), xyz, w0, parent), w0)
    # sum over atoms of w_k(r) = 1: the molecular grid integrates a smooth, decaying function like one big atomic grid
    coords = np.array([[0.0, 0.0, -1.1], [0.0, 0.0, 1.1], [1.5, 0.3, 0.0]])
    for flavour in ("SSF", "BECKE"):
        xyz, w0, parent = reference_atom_grids(["O", "C", "H"], coords, 4)
        w = host_partition_weights(flavour, [8, 6, 1], coords, xyz, w0, parent)
        f = sum(np.exp(-0.7 * ((xyz - c) ** 2).sum(axis=1)) for c in coords)
        assert abs((w * f).sum() - 3.0 * (math.pi / 0.7) ** 1.5) < 2e-4  # acc-4 quadrature error, 28.5 in total


def _compare(ctx, symbols, coords, acc, flavour, smoothing=3):
    from serenity_b200.inputs.geometry import atomic_numbers
    from serenity_b200.inputs.grid import becke_size_adjustments, host_partition_weights, reference_atom_grids
    zs = atomic_numbers(symbols)
    xyz, w0, parent = reference_atom_grids(symbols, coords, acc)
    ref = host_partition_weights(flavour, zs, coords, xyz, w0, parent, smoothing)
    got, ms = ctx.partition_weights(flavour, coords, xyz, parent, w0,
                                    becke_size_adjustments(zs) if flavour != "SSF" else None, smoothing)
    # the cell functions are the same products in the same order; the sum over atoms is reduced in a different order
    # and FMA contraction may differ -> relative 1e-12, absolute 1e-15 of the largest atomic weight
    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref) + 1e-15 * np.abs(w0).max()), float(np.abs(got - ref).max())
    # identical set of surviving points, except where a weight sits on the cut itself
    border = np.abs(ref - 1e-14) < 1e-20
    assert np.array_equal((got > 1e-14)[~border], (ref > 1e-14)[~border])
    return xyz, got, ms


@pytest.mark.gpu
def test_gpu_partition_weights_match_host_and_reference_kats():
    from serenity_b200.xc import XCContext
    c60, radius, cases = _kats()
    ctx = XCContext(0)
    for c in cases:
        xyz, w, _ = _compare(ctx, ["C"] * 60, c60, c["accuracy"], c["flavour"], c["smoothing"])
        keep = w > c["weight_threshold"]
        _check_window(c, _sphere_ratio(xyz[keep], w[keep], radius))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name,acc", [("h2o", 4), ("tetracene", 4), ("water64", 2)])
def test_gpu_partition_weights_on_the_baseline_geometries(name, acc):
    from serenity_b200.inputs.configs import geometry_of
    from serenity_b200.xc import XCContext
    symbols, coords = geometry_of(name)
    ctx = XCContext(0)
    _compare(ctx, symbols, coords, acc, "SSF")
    if name == "h2o":
        _compare(ctx, symbols, coords, acc, "BECKE")
        one = np.zeros((1, 3))
        from serenity_b200.inputs.grid import reference_atom_grids
        xyz, w0, parent = reference_atom_grids(["O"], one, 2)
        got, _ = ctx.partition_weights("SSF", one, xyz, parent, w0)
        assert np.array_equal(got, w0)  # GridFactory.cpp: a single atom keeps its atomic grid
    ctx.close()


@pytest.mark.gpu
def test_gpu_grid_built_on_device_gives_the_same_xc_build():
    """molecular_grid(device_ctx=...) is a drop-in for the host producer: same points, E_xc to 1e-12."""
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.inputs.grid import molecular_grid
    from serenity_b200.xc import XCContext
    cfg = make_config("h2o", 4)
    sub = cfg.subsystems[0]
    ctx = XCContext(0)
    xyz, w = molecular_grid(sub.symbols, sub.coords, 4, device_ctx=ctx)
    assert xyz.shape == cfg.xyz.shape and np.array_equal(xyz, cfg.xyz)
    assert np.abs(w - cfg.w).max() <= 1e-12 * np.abs(cfg.w).max()
    f = ctx.set_functional(*FUNCTIONALS["PBE"])
    b = ctx.add_basis(sub.basis, 1e-9)
    E = [ctx.build_xc(ctx.set_grid(x, ww, 128), b, f, sub.P)[1] for x, ww in ((cfg.xyz, cfg.w), (xyz, w))]
    assert abs(E[0] - E[1]) < 1e-12
    ctx.close()


def test_atom_grid_factory_property_tests_of_the_reference():
    """AtomGridFactory_test.cpp:46-75 (CheckRadialPointsByIntegration) and :80-99 (CheckSphericalPointsByIntegration) replayed on
    the host restatement of the atom grids (serenity_b200/inputs/grid.py): the Ahlrichs radial rule integrates a Gaussian of
    width c to 1/2 sqrt(2 pi) c within 1e-6 for every (c, nRad, alpha) of the reference's loops; the Lebedev rules of the
    first 20 indices have unit-norm points and weights summing to 1 within 1e-8."""
    import math
    from serenity_b200.inputs.grid import _lebedev, ahlrichs_radial
    c = 5.0
    while c > 0.1:
        n_rad = 3000
        while n_rad <= 100000:
            alpha = 0.8
            while alpha <= 2.6:
                r, w = ahlrichs_radial(alpha, n_rad)
                integral = float(np.sum(np.exp(-(r[1:] ** 2) / (2 * c * c)) * w[1:] / r[1:] ** 2))
                assert abs(integral - 0.5 * math.sqrt(2.0 * math.pi) * c) < 1e-6, (c, n_rad, alpha, integral)
                alpha += 0.8
            n_rad = int(n_rad * 6.37)
        c *= 0.2
    for index in range(20):
        x, w = _lebedev(index)
        assert x.shape[0] == w.shape[0]
        assert np.abs(np.sum(x * x, axis=1) - 1.0).max() < 1e-8 and abs(w.sum() - 1.0) < 1e-8


def test_hilbert_rtree_sort_known_answer_of_the_reference():
    """HilbertRTreeSorting_test.cpp:32-54 (2x2x2): eight cube corners with weights 0..7 come out in the order 3 7 1 4 6 2 5 0."""
    from serenity_b200.inputs.grid import hilbert_rtree_order, molecular_grid
    # Eigen's comma initialiser fills the 3 x 8 matrix row by row: all x, then all y, then all z
    x = [+0.5, -0.5, +0.5, -0.5, -0.5, +0.5, +0.5, -0.5]
    y = [+0.5, -0.5, -0.5, +0.5, +0.5, -0.5, +0.5, -0.5]
    z = [+0.5, -0.5, -0.5, +0.5, -0.5, +0.5, -0.5, +0.5]
    xyz = np.stack([x, y, z], axis=1)
    order = hilbert_rtree_order(xyz)
    assert list(order) == [3, 7, 1, 4, 6, 2, 5, 0]
    px = [-0.5, -0.5, -0.5, -0.5, +0.5, +0.5, +0.5, +0.5]
    py = [+0.5, -0.5, -0.5, +0.5, +0.5, -0.5, -0.5, +0.5]
    pz = [+0.5, +0.5, -0.5, -0.5, -0.5, -0.5, +0.5, +0.5]
    assert np.array_equal(xyz[order], np.stack([px, py, pz], axis=1))
    # the sort is a permutation that keeps neighbours together: on a molecular grid consecutive points are close
    gx, gw = molecular_grid(["O", "H", "H"], np.array([[0.0, 0.0, 0.2], [0.0, 1.4, -0.9], [0.0, -1.4, -0.9]]), 2, sort="reference")
    gu, wu = molecular_grid(["O", "H", "H"], np.array([[0.0, 0.0, 0.2], [0.0, 1.4, -0.9], [0.0, -1.4, -0.9]]), 2, sort=False)
    assert gx.shape == gu.shape and abs(gw.sum() - wu.sum()) < 1e-12 * abs(wu.sum())
    step_sorted = np.linalg.norm(np.diff(gx, axis=0), axis=1).mean()
    step_unsorted = np.linalg.norm(np.diff(gu, axis=0), axis=1).mean()
    assert step_sorted < 0.5 * step_unsorted


# ------------------------------------------------------------------------------------------------ grid construction in C++ (f-1)
@pytest.mark.parametrize("z,acc,radial", [(1, 4, "AHLRICHS"), (6, 4, "AHLRICHS"), (8, 6, "AHLRICHS"), (7, 2, "AHLRICHS"), (1, 1, "AHLRICHS"),
                                          (6, 4, "BECKE"), (1, 3, "BECKE")])
def test_cpp_atom_grid_matches_the_python_restatement(z, acc, radial):
    """sxc_atom_grid (csrc/grid_builder.cpp: AtomGridFactory.cpp:78-255) against serenity_b200.inputs.grid.atom_grid, which meets
    the property tests of AtomGridFactory_test.cpp:46-99: same point count and order, coordinates and weights to rounding."""
    from serenity_b200.inputs.grid import atom_grid as py_atom_grid
    from serenity_b200.xc import atom_grid
    xyz, w = atom_grid(z, acc, radial)
    pxyz, pw = py_atom_grid(z, acc, radial)
    assert xyz.shape == pxyz.shape and w.shape == pw.shape
    assert np.abs(xyz - pxyz).max() <= 1e-13 * np.abs(pxyz).max()
    assert np.abs(w - pw).max() <= 1e-13 * np.abs(pw).max()
    # the unit sphere integrates to its volume through every shell structure (Lebedev weights sum to 1 per shell)
    r = np.linalg.norm(xyz, axis=1)
    f = np.exp(-r * r)
    assert abs(np.dot(w, f) - np.pi ** 1.5) < 2e-5 * np.pi ** 1.5


def test_cpp_atom_grid_rejects_what_it_has_no_data_for():
    from serenity_b200._lib import SerenityError
    from serenity_b200.xc import atom_grid
    with pytest.raises(SerenityError, match="H..Kr"):
        atom_grid(40, 4)
    with pytest.raises(SerenityError, match="Bragg-Slater"):
        atom_grid(2, 4, "BECKE")


def test_cpp_hilbert_rtree_order_reproduces_the_reference_known_answer():
    """HilbertRTreeSorting_test.cpp:32-54 (the 2 x 2 x 2 cube) through sxc_hilbert_rtree_order, and the Python restatement on a
    molecular point cloud."""
    from serenity_b200.inputs.grid import hilbert_rtree_order as py_order
    from serenity_b200.xc import hilbert_rtree_order
    x = [+0.5, -0.5, +0.5, -0.5, -0.5, +0.5, +0.5, -0.5]  # (Eigen's comma initialiser fills the 3 x 8 matrix row by row)
    y = [+0.5, -0.5, -0.5, +0.5, +0.5, -0.5, +0.5, -0.5]
    z = [+0.5, -0.5, -0.5, +0.5, -0.5, +0.5, -0.5, +0.5]
    assert list(hilbert_rtree_order(np.stack([x, y, z], axis=1))) == [3, 7, 1, 4, 6, 2, 5, 0]
    rng = np.random.default_rng(5)
    cloud = rng.normal(size=(20000, 3)) * np.array([3.0, 1.0, 2.0])
    assert np.array_equal(hilbert_rtree_order(cloud), py_order(cloud))


@pytest.mark.gpu
@pytest.mark.parametrize("flavour", ["SSF", "BECKE"])
def test_cpp_molecular_grid_matches_the_python_pipeline(flavour):
    """GridFactory::produce behind the C ABI (sxc_molecular_grid): same points in the same order as the Python pipeline with the
    reference's sort, weights to rounding."""
    from serenity_b200.inputs import geometry as geo
    from serenity_b200.inputs.geometry import atomic_numbers
    from serenity_b200.inputs.grid import molecular_grid
    from serenity_b200.xc import XCContext
    symbols, coords = geo.water_cluster(2)
    ctx = XCContext(0)
    try:
        xyz, w = ctx.molecular_grid(atomic_numbers(symbols), coords, 3, flavour)
        pxyz, pw = molecular_grid(symbols, coords, 3, flavour, sort="reference", device_ctx=ctx)
        assert xyz.shape == pxyz.shape
        assert np.abs(xyz - pxyz).max() <= 1e-12 and np.abs(w - pw).max() <= 1e-12 * np.abs(pw).max()
    finally:
        ctx.close()
