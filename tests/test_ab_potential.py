"""Two-basis operators: ScalarOperatorToMatrixAdder with basis A != basis B
(src/data/grid/ScalarOperatorToMatrixAdder.cpp:216-220, :286-300) and ABFuncPotential::getMatrix
(src/potentials/ABFockMatrixConstruction/ABFuncPotential.cpp:54-160).

Known answers: the reference's ABFuncPotential_test.cpp:108-172 evaluates the A x B matrix of H2 / def2-TZVP (system A, its
grid and stored density matrix) against the 6-31G* basis of the H2 environment molecule (system B) and pins six elements
for LDA (restricted) and for BP86 (unrestricted, alpha = beta) to 1e-5; ABFuncPotential_test.cpp:58-106 demands that
A x A reproduces FuncPotential to 1e-12.  The inputs are the ones of tests/golden/h2_kats.json.
"""
import numpy as np
import pytest

from conftest import load_golden
from test_reference_kats import BP86, LDA, _grid, _system

# ABFuncPotential_test.cpp:123-128 (LDA, RESTRICTED) and :150-155 (BP86, UNRESTRICTED alpha == beta), tolerance 1e-5
AB_LDA = [((0, 0), -3.636117713288e-02), ((1, 0), -1.001498923041e-01), ((2, 0), -1.061975191271e-01),
          ((0, 1), -1.116016285240e-01), ((0, 2), -6.955048960966e-03), ((0, 3), -6.493820327824e-02)]
AB_BP86 = [((0, 0), -3.801470612146e-02), ((1, 0), -1.037127215935e-01), ((2, 0), -1.091671374726e-01),
           ((0, 1), -1.165296890920e-01), ((0, 2), -7.273116668792e-03), ((0, 3), -6.778065332781e-02)]


@pytest.fixture(scope="module")
def h2_ab():
    d = load_golden("h2_kats.json")
    k, kn = d["func_potential"], d["nadd_potential"]
    syms, xyz, tab_a = _system(k["basis_shells_H"], k["geometry_angstrom"], k["settings"]["spherical"])
    gx, gw = _grid(syms, xyz, k["settings"])
    _, _, tab_b = _system(kn["basis_shells_H"], kn["env"]["geometry_angstrom"], kn["env"]["settings"]["spherical"])
    mat = lambda key: np.asarray(k[key]).reshape(12, 12)  # noqa: E731
    return tab_a, tab_b, gx, gw, mat


def _check(V, rows, tol):
    for (i, j), ref in rows:
        assert abs(V[i, j] - ref) < tol, (i, j, V[i, j], ref)


def test_oracle_ab_reference_kats(h2_ab):
    from oracle import pyoracle as orc
    tab_a, tab_b, gx, gw, mat = h2_ab
    ba, bb, og = orc.Basis(tab_a), orc.Basis(tab_b), orc.Grid(gx, gw, 128)
    P = mat("P_restricted")
    V, _ = orc.build_ab(ba, bb, [(ba, P)], og, orc.Functional(*LDA))
    assert V.shape == (12, 4)
    _check(V, AB_LDA, 1e-7)    # reference tolerance 1e-5; measured 6e-10
    V, _ = orc.build_ab(ba, bb, [(ba, P)], og, orc.Functional(*BP86))
    _check(V, AB_BP86, 1e-7)
    # A x A is the symmetric path (ABFuncPotential_test.cpp:58-83: "allow only white noise")
    Vaa, E = orc.build_ab(ba, ba, [(ba, P)], og, orc.Functional(*BP86))
    Vs, Es, _, _ = orc.build_xc(ba, og, orc.Functional(*BP86), P)
    assert np.abs(Vaa - Vs).max() < 1e-12 and abs(E - Es) < 1e-12


@pytest.mark.gpu
def test_gpu_ab_reference_kats(h2_ab):
    from serenity_b200.xc import XCContext
    tab_a, tab_b, gx, gw, mat = h2_ab
    ctx = XCContext(0)
    g = ctx.set_grid(gx, gw, 128)
    ba, bb = ctx.add_basis(tab_a, 1e-9), ctx.add_basis(tab_b, 1e-9)
    P = mat("P_restricted")
    V, _, ne = ctx.build_ab(g, ctx.set_functional(*LDA), ba, bb, 12, 4, [ba], [P])
    _check(V, AB_LDA, 1e-7)
    assert abs(ne - 2.0) < 1e-6
    fb = ctx.set_functional(*BP86)
    (Va, Vb), _, _ = ctx.build_ab(g, fb, ba, bb, 12, 4, [ba], [(mat("P_alpha"), mat("P_beta"))], nspin=2)
    _check(Va, AB_BP86, 1e-7)
    _check(Vb, AB_BP86, 1e-7)
    Vaa, E, _ = ctx.build_ab(g, fb, ba, ba, 12, 12, [ba], [P])
    Vs, Es, _ = ctx.build_xc(g, ba, fb, P)
    assert np.abs(Vaa - Vs).max() < 1e-12 and abs(E - Es) < 1e-12
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("func", ["LDA", "PBE", "B3LYP"])
def test_gpu_ab_matches_oracle_on_the_water_dimer(func):
    """A = first water (24 functions), B = second water, density of both molecules, supersystem grid."""
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    cfg = make_config("fde_dimer", 2)
    sa, sb = cfg.subsystems
    ids, mix = FUNCTIONALS[func]
    oa, ob, og = orc.Basis(sa.basis), orc.Basis(sb.basis), orc.Grid(cfg.xyz, cfg.w, 128)
    V_ref, E_ref = orc.build_ab(oa, ob, [(oa, sa.P), (ob, sb.P)], og, orc.Functional(ids, mix))
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, bb = ctx.add_basis(sa.basis, 1e-9), ctx.add_basis(sb.basis, 1e-9)
    f = ctx.set_functional(ids, mix)
    na, nb = sa.basis.nbf, sb.basis.nbf
    V, E, _ = ctx.build_ab(g, f, ba, bb, na, nb, [ba, bb], [sa.P, sb.P])
    assert np.abs(V - V_ref).max() <= 1e-8 and abs(E - E_ref) <= 1e-9       # north_star tolerances
    # closed-shell consistency of the UNRESTRICTED path
    (Va, Vb), Eu, _ = ctx.build_ab(g, f, ba, bb, na, nb, [ba, bb], [(0.5 * sa.P, 0.5 * sa.P), (0.5 * sb.P, 0.5 * sb.P)], nspin=2)
    assert np.abs(Va - V).max() < 1e-10 and np.abs(Vb - V).max() < 1e-10 and abs(Eu - E) < 1e-10
    # stage-level entry: an arbitrary operator on the grid, added into the caller's matrix
    rng = np.random.default_rng(3)
    v, gx, gy, gz = (rng.standard_normal(cfg.npts) * 1e-2 for _ in range(4))
    W_ref = orc.scalar_to_matrix_ab(oa, ob, og, 1e-9, 1e-11, v, gx, gy, gz)
    W = ctx.scalar_to_matrix_ab(g, ba, bb, na, nb, v, gx, gy, gz, V=np.ones((na, nb), order="F"))
    assert np.abs(W - 1.0 - W_ref).max() <= 1e-10
    W_ref = orc.scalar_to_matrix_ab(oa, ob, og, 1e-9, 1e-11, v)
    W = ctx.scalar_to_matrix_ab(g, ba, bb, na, nb, v)
    assert np.abs(W - W_ref).max() <= 1e-10
    ctx.close()


@pytest.mark.gpu
def test_gpu_ab_rectangular_large():
    """s_A and s_B of a few hundred functions: several bands / column chunks per block (tetracene x a water cluster basis)."""
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.basis import build_shell_table
    from serenity_b200.inputs.configs import FUNCTIONALS, geometry_of
    from serenity_b200.xc import XCContext
    cfg = make_config("tetracene", 2)
    sa = cfg.subsystems[0]
    sym_b, xyz_b = geometry_of("water8")
    tab_b = build_shell_table(sym_b, xyz_b * 0.6, "def2-svp", spherical=True)  # squeezed into the tetracene's volume
    ids, mix = FUNCTIONALS["PBE"]
    oa, ob, og = orc.Basis(sa.basis), orc.Basis(tab_b), orc.Grid(cfg.xyz, cfg.w, 128)
    V_ref, E_ref = orc.build_ab(oa, ob, [(oa, sa.P)], og, orc.Functional(ids, mix))
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, bb = ctx.add_basis(sa.basis, 1e-9), ctx.add_basis(tab_b, 1e-9)
    V, E, _ = ctx.build_ab(g, ctx.set_functional(ids, mix), ba, bb, sa.basis.nbf, tab_b.nbf, [ba], [sa.P])
    assert np.abs(V - V_ref).max() <= 1e-8 and abs(E - E_ref) <= 1e-9
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------------
# ABNAddFuncPotential (potentials/ABFockMatrixConstruction/ABNAddFuncPotential.cpp:66-176): known answers of
# ABNAddFuncPotential_test.cpp:86-151 - H2 (6-31G*, active FDE system, its SSF grid) against the basis of the environment H2,
# LDA / BP86, one and two environment density matrices; tolerance of the reference 1e-5.
ABN_LDA = [((0, 0), -0.0083840258804812658), ((1, 0), -0.055300758004326538), ((2, 0), -0.12831003302437624),
           ((0, 1), -0.005639132369059649), ((0, 2), -0.0039221790739654878), ((0, 3), -0.0044904816299493342)]
ABN_BP86 = [((0, 0), -0.0076542762861270645), ((1, 0), -0.055253744527196447), ((2, 0), -0.12698015387994177),
            ((0, 1), -0.0047051305109973322), ((0, 2), -0.0037534727526342374), ((0, 3), -0.0038349418254594093)]
ABN_BP86_2ENV = [((0, 0), -0.012352511055803322), ((1, 0), -0.080462987092342453), ((2, 0), -0.18794150046341079),
                 ((0, 1), -0.0081584035257395351), ((0, 2), -0.0057641379450071991), ((0, 3), -0.0065285874827300102)]


@pytest.fixture(scope="module")
def h2_fde():
    k = load_golden("h2_kats.json")["nadd_potential"]
    st = k["settings"]
    syms, xyz, tab_a = _system(k["basis_shells_H"], k["act"]["geometry_angstrom"], st["spherical"])
    _, _, tab_b = _system(k["basis_shells_H"], k["env"]["geometry_angstrom"], st["spherical"])
    gx, gw = _grid(syms, xyz, st)  # the active system's grid (ABNAddFuncPotential_test.cpp:65)
    P = lambda s: np.asarray(k[s]["P_restricted"]).reshape(4, 4)  # noqa: E731
    return tab_a, tab_b, gx, gw, P("act"), P("env")


def test_oracle_abnadd_reference_kats(h2_fde):
    from oracle import pyoracle as orc
    tab_a, tab_b, gx, gw, Pa, Pe = h2_fde
    ba, bb, og = orc.Basis(tab_a), orc.Basis(tab_b), orc.Grid(gx, gw, 128)
    _check(orc.build_ab_nadd(ba, bb, (ba, Pa), [(bb, Pe)], og, orc.Functional(*LDA)), ABN_LDA, 1e-6)
    _check(orc.build_ab_nadd(ba, bb, (ba, Pa), [(bb, Pe)], og, orc.Functional(*BP86)), ABN_BP86, 1e-6)
    _check(orc.build_ab_nadd(ba, bb, (ba, Pa), [(bb, Pe), (bb, Pe)], og, orc.Functional(*BP86)), ABN_BP86_2ENV, 1e-6)
    # A x A is NAddFuncPotential (:59-80, "accept only white noise")
    V_aa = orc.build_ab_nadd(ba, ba, (ba, Pa), [(bb, Pe)], og, orc.Functional(*LDA))
    V_n, _, _ = orc.build_nadd(ba, Pa, [(bb, Pe)], og, orc.Functional(*LDA))
    assert np.abs(V_aa - V_n).max() < 1e-12


@pytest.mark.gpu
def test_gpu_abnadd_reference_kats(h2_fde):
    from serenity_b200.xc import XCContext
    tab_a, tab_b, gx, gw, Pa, Pe = h2_fde
    ctx = XCContext(0)
    g = ctx.set_grid(gx, gw, 128)
    ba, bb = ctx.add_basis(tab_a, 1e-9), ctx.add_basis(tab_b, 1e-9)
    fl, fb = ctx.set_functional(*LDA), ctx.set_functional(*BP86)
    _check(ctx.build_ab_nadd(g, fl, ba, bb, 4, 4, ba, Pa, [bb], [Pe]), ABN_LDA, 1e-6)
    _check(ctx.build_ab_nadd(g, fb, ba, bb, 4, 4, ba, Pa, [bb], [Pe]), ABN_BP86, 1e-6)
    _check(ctx.build_ab_nadd(g, fb, ba, bb, 4, 4, ba, Pa, [bb, bb], [Pe, Pe]), ABN_BP86_2ENV, 1e-6)
    V_aa = ctx.build_ab_nadd(g, fl, ba, ba, 4, 4, ba, Pa, [bb], [Pe])
    V_n, _ = ctx.build_nadd(g, fl, ba, Pa, [bb], [Pe])
    assert np.abs(V_aa - V_n).max() < 1e-12
    # closed-shell consistency of the UNRESTRICTED build
    Va, Vb = ctx.build_ab_nadd(g, fb, ba, bb, 4, 4, ba, (0.5 * Pa, 0.5 * Pa), [bb], [(0.5 * Pe, 0.5 * Pe)], nspin=2)
    Vr = ctx.build_ab_nadd(g, fb, ba, bb, 4, 4, ba, Pa, [bb], [Pe])
    assert np.abs(Va - Vr).max() < 1e-12 and np.abs(Vb - Vr).max() < 1e-12
    ctx.close()


@pytest.mark.gpu
def test_gpu_abnadd_matches_oracle_on_the_water_dimer():
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    cfg = make_config("fde_dimer", 2)
    sa, sb = cfg.subsystems
    oa, ob, og = orc.Basis(sa.basis), orc.Basis(sb.basis), orc.Grid(cfg.xyz, cfg.w, 128)
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    ba, bb = ctx.add_basis(sa.basis, 1e-9), ctx.add_basis(sb.basis, 1e-9)
    for name in ("PBE", "PW91K", "LDA"):
        ids, mix = FUNCTIONALS[name]
        V_ref = orc.build_ab_nadd(oa, ob, (oa, sa.P), [(ob, sb.P)], og, orc.Functional(ids, mix))
        V = ctx.build_ab_nadd(g, ctx.set_functional(ids, mix), ba, bb, sa.basis.nbf, sb.basis.nbf, ba, sa.P, [bb], [sb.P])
        assert np.abs(V - V_ref).max() <= 1e-8, (name, np.abs(V - V_ref).max())
    ctx.close()
