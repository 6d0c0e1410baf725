"""XC nuclear gradient: FuncPotential::getGeomGradients (src/potentials/FuncPotential.cpp:114-239).

The reference's gradient KATs (FuncPotential_test.cpp:234-330) need a converged SCF density built from scratch and
cannot be replayed offline; the oracle's restatement of the double loop is therefore pinned to the DEFINITION instead:
with P and the grid held fixed, g[A] equals the derivative of E_xc with respect to a rigid shift of the basis functions
centred on atom A (the reference neglects grid-weight derivatives in the same way).  The CUDA path is then compared with
the oracle.
"""
import copy

import numpy as np
import pytest


def _case(name, acc, func):
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.basis import atom_indices_of_basis
    from serenity_b200.inputs.configs import FUNCTIONALS
    cfg = make_config(name, acc)
    sub = cfg.subsystems[0]
    return cfg, sub, FUNCTIONALS[func], atom_indices_of_basis(sub.basis, sub.coords), len(sub.symbols)


def _shifted(tab, amap, atom, d):
    t = copy.copy(tab)
    t.centre = tab.centre.copy()
    for sh in range(tab.nshell):
        if amap[tab.first_bf[sh]] == atom:
            t.centre[sh] += d
    return t


@pytest.mark.parametrize("func", ["LDA", "PBE"])
def test_oracle_gradient_is_derivative_of_exc(func):
    from oracle import pyoracle as orc
    cfg, sub, (ids, mix), amap, natoms = _case("h2o", 2, func)
    og, of = orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix)
    g = orc.xc_gradient(orc.Basis(sub.basis), og, of, sub.P, amap, natoms)
    h = 1e-4
    for atom, c in ((0, 2), (1, 0), (2, 1)):
        d = np.zeros(3)
        d[c] = h
        ep = orc.build_xc(orc.Basis(_shifted(sub.basis, amap, atom, d)), og, of, sub.P)[1]
        em = orc.build_xc(orc.Basis(_shifted(sub.basis, amap, atom, -d)), og, of, sub.P)[1]
        fd = (ep - em) / (2 * h)
        assert abs(fd - g[atom, c]) < 2e-7 * max(1.0, abs(fd)), (atom, c, fd, g[atom, c])


def test_oracle_gradient_unrestricted_reduces_to_restricted():
    from oracle import pyoracle as orc
    cfg, sub, (ids, mix), amap, natoms = _case("h2o", 2, "PBE")
    ob, og, of = orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix)
    g = orc.xc_gradient(ob, og, of, sub.P, amap, natoms)
    gu = orc.xc_gradient(ob, og, of, (0.5 * sub.P, 0.5 * sub.P), amap, natoms)
    assert np.abs(g - gu).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name,acc,func", [("h2o", 4, "PBE"), ("h2o", 2, "LDA"), ("h2o", 2, "B3LYP"), ("water8", 2, "PBE")])
def test_gpu_gradient_matches_oracle(name, acc, func):
    from oracle import pyoracle as orc
    from serenity_b200.xc import XCContext
    cfg, sub, (ids, mix), amap, natoms = _case(name, acc, func)
    ref = orc.xc_gradient(orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix), sub.P, amap, natoms)
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    b = ctx.add_basis(sub.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    got = ctx.xc_gradient(g, b, f, sub.P, amap, natoms)
    assert np.abs(got - ref).max() < 1e-9, np.abs(got - ref).max()
    # a potential build afterwards still works on the same handles (separate 8-slot plan)
    V, E, _ = ctx.build_xc(g, b, f, sub.P)
    assert np.isfinite(V).all()
    ctx.close()


@pytest.mark.gpu
def test_gpu_gradient_unrestricted_and_cartesian():
    from oracle import pyoracle as orc
    from serenity_b200.inputs.basis import atom_indices_of_basis, build_shell_table
    from serenity_b200.xc import XCContext
    cfg, sub, (ids, mix), amap, natoms = _case("h2o", 2, "PBE")
    rng = np.random.default_rng(5)
    nb, nocc = sub.basis.nbf, 5
    C = rng.normal(size=(nb, nocc + 1)) / np.sqrt(nb)
    Pa, Pb = np.asfortranarray(C @ C.T), np.asfortranarray(C[:, :nocc - 1] @ C[:, :nocc - 1].T)
    og, of = orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(ids, mix)
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    f = ctx.set_functional(ids, mix)
    b = ctx.add_basis(sub.basis, 1e-9)
    ref = orc.xc_gradient(orc.Basis(sub.basis), og, of, (Pa, Pb), amap, natoms)
    got = ctx.xc_gradient(g, b, f, (Pa, Pb), amap, natoms, nspin=2)
    assert np.abs(got - ref).max() < 1e-9
    # Cartesian shells (d functions of def2-SVP on oxygen as 6 Cartesians)
    tabc = build_shell_table(sub.symbols, sub.coords, "def2-svp", spherical=False)
    amapc = atom_indices_of_basis(tabc, sub.coords)
    Cc = rng.normal(size=(tabc.nbf, nocc)) / np.sqrt(tabc.nbf)
    Pc = np.asfortranarray(2 * Cc @ Cc.T)
    bc = ctx.add_basis(tabc, 1e-9)
    refc = orc.xc_gradient(orc.Basis(tabc), og, of, Pc, amapc, natoms)
    gotc = ctx.xc_gradient(g, bc, f, Pc, amapc, natoms)
    assert np.abs(gotc - refc).max() < 1e-9
    ctx.close()
