"""Known-answer tests at the Potential level replayed from the reference's own tests (tests/golden/h2_kats.json, made by
tests/golden/make_h2_kats.py): FuncPotential_test.cpp:59-228 (12 x 12 V_xc of H2/def2-TZVP, LDA and BP86, R and U)
and NAddFuncPotential_test.cpp:58-185 (V_nadd elements of the H2 dimer, LDA and BP86).

These pin the WHOLE path - grid restatement (Becke / SSF partition, Ahlrichs radial, pruned Lebedev), libint-style
renormalisation, Cartesian shells, density, functional kernels, grid -> matrix - against numbers produced by the
reference itself (xcfun route):
  LDA  = slaterx + vwn5c : agrees to 5e-9 (reference tolerance 1e-6)           -> Slater and VWN5 kernels PINNED
  BP86 = beckex + p86c   : agrees to 8.4e-8 (reference tolerance 1e-6)        -> B88, PZ81 and P86 kernels PINNED.
                           (With the paper's rounded prefactor 1.745 in the P86 exponent the deviation was 8e-6; a
                           one-parameter fit of the reference matrix gave 1.000238 = (9 pi)^(1/6) / 1.745, i.e. xcfun
                           uses the exact constant; all other parameters then fit to 1 +- 2e-7)
  NAdd LDA / BP86 elements (RESTRICTED): inside the reference tolerance 1e-5.
  E_xc[BP86] stored in data/testresources/TestSystem_H2_6_31Gs_{ACTIVE,ENVIRONMENT}_FDE/*.energies.res (6 decimals):
                           reproduced to 4e-7 from the density matrices stored next to them.
"""
import numpy as np
import pytest

from conftest import load_golden

LDA = ([2, 45], [1.0, 1.0])     # CompositeFunctionals.cpp:233-235
BP86 = ([80, 193], [1.0, 1.0])  # CompositeFunctionals.cpp:244-246


def _system(shells_h, geometry, spherical):
    from serenity_b200.inputs.basis import shell_table_from_list
    from serenity_b200.inputs.geometry import ANGSTROM_TO_BOHR
    xyz = np.asarray([a[1:4] for a in geometry], dtype=np.float64) * ANGSTROM_TO_BOHR
    shells = [{"l": sh["l"], "pure": spherical, "exps": sh["exps"], "coefs": sh["coefs"], "centre": list(c)}
              for c in xyz for sh in shells_h]
    return [a[0] for a in geometry], xyz, shell_table_from_list(shells)


def _grid(syms, xyz, st):
    from serenity_b200.inputs.grid import molecular_grid
    return molecular_grid(syms, xyz, st["accuracy"], flavour=st["grid_type"], radial=st["radial"],
                          weight_threshold=st["weight_threshold"])


@pytest.fixture(scope="module")
def func_case():
    k = load_golden("h2_kats.json")["func_potential"]
    syms, xyz, tab = _system(k["basis_shells_H"], k["geometry_angstrom"], k["settings"]["spherical"])
    gx, gw = _grid(syms, xyz, k["settings"])
    mat = lambda key: np.asarray(k[key]).reshape(12, 12)  # noqa: E731
    return k, tab, gx, gw, mat


@pytest.fixture(scope="module")
def nadd_case():
    k = load_golden("h2_kats.json")["nadd_potential"]
    st = k["settings"]
    syms, xyz, tabA = _system(k["basis_shells_H"], k["act"]["geometry_angstrom"], st["spherical"])
    _, _, tabE = _system(k["basis_shells_H"], k["env"]["geometry_angstrom"], st["spherical"])
    gx, gw = _grid(syms, xyz, st)  # the test integrates on the ACTIVE system's grid (NAddFuncPotential_test.cpp:68)
    return k, tabA, tabE, gx, gw


# ------------------------------------------------------------------------------------------- oracle (CPU)
def test_oracle_h2_vxc_lda_pinned(func_case):
    from oracle import pyoracle as orc
    k, tab, gx, gw, mat = func_case
    ob, og = orc.Basis(tab), orc.Grid(gx, gw, 128)
    V, E, ne, _ = orc.build_xc(ob, og, orc.Functional(*LDA), mat("P_restricted"), k["settings"]["radial_threshold"],
                               k["settings"]["block_ave_threshold"])
    assert abs(ne - 2.0) < 1e-6
    assert np.abs(V - mat("V_LDA")).max() < 2e-8          # reference tolerance 1e-6; the table has 8 decimals
    (Va, Vb), Eu, _ = orc.build_xc_u(ob, og, orc.Functional(*LDA), mat("P_alpha"), mat("P_beta"))
    assert np.abs(Va - mat("V_LDA_unres")).max() < 3e-8 and np.abs(Vb - mat("V_LDA_unres")).max() < 3e-8


def test_oracle_h2_vxc_bp86(func_case):
    from oracle import pyoracle as orc
    k, tab, gx, gw, mat = func_case
    ob, og = orc.Basis(tab), orc.Grid(gx, gw, 128)
    V, _, _, _ = orc.build_xc(ob, og, orc.Functional(*BP86), mat("P_restricted"))
    dev = np.abs(V - mat("V_BP86")).max()
    assert dev < 2e-7, dev          # reference tolerance 1e-6; measured 8.4e-8 (module docstring)
    (Va, Vb), _, _ = orc.build_xc_u(ob, og, orc.Functional(*BP86), mat("P_alpha"), mat("P_beta"))
    assert np.abs(Va - mat("V_BP86_unres")).max() < 2e-7 and np.abs(Vb - mat("V_BP86_unres")).max() < 2e-7


def _check_elements(V, rows):
    for i, j, ref, tol in rows:
        assert abs(V[i, j] - ref) < tol, (i, j, V[i, j], ref)


def test_oracle_h2_dimer_nadd(nadd_case):
    from oracle import pyoracle as orc
    k, tabA, tabE, gx, gw = nadd_case
    bA, bE, og = orc.Basis(tabA), orc.Basis(tabE), orc.Grid(gx, gw, 128)
    P = lambda s, key: np.asarray(k[s][key]).reshape(4, 4)  # noqa: E731
    for name, fn in (("LDA", LDA), ("BP86", BP86)):
        V, _, _ = orc.build_nadd(bA, P("act", "P_restricted"), [(bE, P("env", "P_restricted"))], og, orc.Functional(*fn))
        _check_elements(V, k[name])
        # closed-shell consistency of the UNRESTRICTED path on the same densities.  (The reference's *_UNRES expectations
        # are not replayed: its unrestricted electronic structure is rebuilt from the stored orbitals, and the
        # .dmat.unres.h5 matrices do not reproduce them - R with P = 2 P_alpha gives -0.003094 vs the expected -0.003314.)
        Ph = (0.5 * P("act", "P_restricted"), 0.5 * P("act", "P_restricted"))
        Pe = (0.5 * P("env", "P_restricted"), 0.5 * P("env", "P_restricted"))
        (Va, Vb), _, _ = orc.build_nadd_u(bA, Ph, [(bE, Pe)], og, orc.Functional(*fn))
        assert np.abs(Va - V).max() < 1e-10 and np.abs(Vb - V).max() < 1e-10


def _stored_energy_cases():
    """(system entry, shell table, grid) of the two H2 / 6-31G* FDE test systems whose `.energies.res` is consistent with
    the stored density matrix: E_xc[BP86] printed with 6 decimals (EnergyContributions.h:49, entry 202)."""
    k = load_golden("h2_kats.json")["nadd_potential"]
    for key in ("act", "env"):
        e = k[key]
        syms, xyz, tab = _system(k["basis_shells_H"], e["geometry_angstrom"], e["settings"]["spherical"])
        gx, gw = _grid(syms, xyz, e["settings"])
        yield e, tab, gx, gw


def test_oracle_stored_bp86_energies():
    from oracle import pyoracle as orc
    for e, tab, gx, gw in _stored_energy_cases():
        _, E, _, _ = orc.build_xc(orc.Basis(tab), orc.Grid(gx, gw, 128), orc.Functional(*BP86),
                                  np.asarray(e["P_restricted"]).reshape(4, 4), e["settings"]["radial_threshold"],
                                  e["settings"]["block_ave_threshold"])
        assert abs(E - e["E_xc_BP86_stored"]) < 1e-6, (e["system"], E)   # the file holds 6 decimals


# ------------------------------------------------------------------------------------------- CUDA path
@pytest.mark.gpu
def test_gpu_h2_vxc_kats(func_case):
    from serenity_b200.xc import XCContext
    k, tab, gx, gw, mat = func_case
    ctx = XCContext(0)
    g = ctx.set_grid(gx, gw, 128)
    b = ctx.add_basis(tab, k["settings"]["radial_threshold"])
    f = ctx.set_functional(*LDA)
    V, _, ne = ctx.build_xc(g, b, f, mat("P_restricted"))
    assert np.abs(V - mat("V_LDA")).max() < 2e-8 and abs(ne - 2.0) < 1e-6
    (Va, Vb), _, _ = ctx.build_xc(g, b, f, (mat("P_alpha"), mat("P_beta")), nspin=2)
    assert np.abs(Va - mat("V_LDA_unres")).max() < 3e-8 and np.abs(Vb - mat("V_LDA_unres")).max() < 3e-8
    f = ctx.set_functional(*BP86)
    V, _, _ = ctx.build_xc(g, b, f, mat("P_restricted"))
    assert np.abs(V - mat("V_BP86")).max() < 2e-7   # reference tolerance 1e-6
    (Va, Vb), _, _ = ctx.build_xc(g, b, f, (mat("P_alpha"), mat("P_beta")), nspin=2)
    assert np.abs(Va - mat("V_BP86_unres")).max() < 2e-7 and np.abs(Vb - mat("V_BP86_unres")).max() < 2e-7
    ctx.close()


@pytest.mark.gpu
def test_gpu_stored_bp86_energies():
    from serenity_b200.xc import XCContext
    ctx = XCContext(0)
    f = ctx.set_functional(*BP86)
    for e, tab, gx, gw in _stored_energy_cases():
        g = ctx.set_grid(gx, gw, 128)
        b = ctx.add_basis(tab, e["settings"]["radial_threshold"])
        _, E, _ = ctx.build_xc(g, b, f, np.asarray(e["P_restricted"]).reshape(4, 4),
                               block_ave_threshold=e["settings"]["block_ave_threshold"])
        assert abs(E - e["E_xc_BP86_stored"]) < 1e-6, (e["system"], E)
    ctx.close()


@pytest.mark.gpu
def test_gpu_h2_dimer_nadd_kats(nadd_case):
    from serenity_b200.xc import XCContext
    k, tabA, tabE, gx, gw = nadd_case
    ctx = XCContext(0)
    g = ctx.set_grid(gx, gw, 128)
    bA, bE = ctx.add_basis(tabA, 1e-9), ctx.add_basis(tabE, 1e-9)
    P = lambda s, key: np.asarray(k[s][key]).reshape(4, 4)  # noqa: E731
    for name, fn in (("LDA", LDA), ("BP86", BP86)):
        f = ctx.set_functional(*fn)
        V, _ = ctx.build_nadd(g, f, bA, P("act", "P_restricted"), [bE], [P("env", "P_restricted")])
        _check_elements(V, k[name])
        Ph = (0.5 * P("act", "P_restricted"), 0.5 * P("act", "P_restricted"))
        Pe = (0.5 * P("env", "P_restricted"), 0.5 * P("env", "P_restricted"))
        (Va, Vb), _ = ctx.build_nadd(g, f, bA, Ph, [bE], [Pe], nspin=2)
        assert np.abs(Va - V).max() < 1e-10 and np.abs(Vb - V).max() < 1e-10
    ctx.close()
