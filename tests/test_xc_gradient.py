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


# ---------------------------------------------------------------------------------------------- NAdd gradient
def _nadd_case():
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.basis import atom_indices_of_basis
    cfg = make_config("fde_dimer", 2)
    act, env = cfg.subsystems
    return cfg, act, env, atom_indices_of_basis(act.basis, act.coords), len(act.symbols)


@pytest.mark.parametrize("func", ["PBE", "PW91K"])
def test_oracle_nadd_gradient_is_derivative_of_e_nadd(func):
    """NAddFuncPotential::getGeomGradients (NAddFuncPotential.cpp:329-493) pinned to its definition: with all density
    matrices and the grid fixed, g[A] = dE_nadd / d(rigid shift of the ACTIVE basis functions on atom A)."""
    from oracle import pyoracle as orc
    from serenity_b200.inputs.configs import FUNCTIONALS
    cfg, act, env, amap, natoms = _nadd_case()
    og, of = orc.Grid(cfg.xyz, cfg.w, 128), orc.Functional(*FUNCTIONALS[func])
    bE = orc.Basis(env.basis)
    g = orc.nadd_gradient(orc.Basis(act.basis), act.P, [(bE, env.P)], og, of, amap, natoms)
    h = 1e-4
    for atom, c in ((0, 2), (1, 0), (2, 1)):
        d = np.zeros(3)
        d[c] = h
        ep = orc.build_nadd(orc.Basis(_shifted(act.basis, amap, atom, d)), act.P, [(bE, env.P)], og, of)[1]
        em = orc.build_nadd(orc.Basis(_shifted(act.basis, amap, atom, -d)), act.P, [(bE, env.P)], og, of)[1]
        fd = (ep - em) / (2 * h)
        assert abs(fd - g[atom, c]) < 2e-7 * max(1.0, abs(fd)), (atom, c, fd, g[atom, c])


@pytest.mark.gpu
@pytest.mark.parametrize("func", ["LDA", "PBE", "PW91K"])
def test_gpu_nadd_gradient_matches_oracle(func):
    from oracle import pyoracle as orc
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    cfg, act, env, amap, natoms = _nadd_case()
    ids, mix = FUNCTIONALS[func]
    ref = orc.nadd_gradient(orc.Basis(act.basis), act.P, [(orc.Basis(env.basis), env.P)], orc.Grid(cfg.xyz, cfg.w, 128),
                            orc.Functional(ids, mix), amap, natoms)
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, be = ctx.add_basis(act.basis, 1e-9), ctx.add_basis(env.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    got = ctx.nadd_gradient(g, f, ba, act.P, [be], [env.P], amap, natoms)
    assert np.abs(got - ref).max() <= 1e-9
    gu = ctx.nadd_gradient(g, f, ba, (0.5 * act.P, 0.5 * act.P), [be], [(0.5 * env.P, 0.5 * env.P)], amap, natoms, nspin=2)
    assert np.abs(gu - got).max() <= 1e-10
    # without an environment the entry point degenerates to FuncPotential::getGeomGradients
    assert np.abs(ctx.nadd_gradient(g, f, ba, act.P, [], [], amap, natoms) - ctx.xc_gradient(g, ba, f, act.P, amap, natoms)).max() < 1e-12
    V1, E1 = ctx.build_nadd(g, f, ba, act.P, [be], [env.P])  # the NAdd build still works after the gradient reused its buffers
    V_ref, E_ref, _ = orc.build_nadd(orc.Basis(act.basis), act.P, [(orc.Basis(env.basis), env.P)], orc.Grid(cfg.xyz, cfg.w, 128),
                                     orc.Functional(ids, mix))
    assert np.abs(V1 - V_ref).max() <= 1e-8
    ctx.close()
