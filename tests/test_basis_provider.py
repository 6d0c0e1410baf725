"""Row f-2: the C++ basis-set front end behind sxc_shell_table_from_file (serenity_b200/csrc/basis_provider.cpp) -
BasisFunctionProvider.cpp:32-140 parsing, Shell.cpp:29-47 / libint2 renormalisation, extended indices.

Checked against (a) the independent Python producer the parity tests have used all along (serenity_b200/inputs/basis.py, whose
renormalisation constants are pinned by the reference's BasisFunctionOnGridController_test vectors, tests/golden/), on
Turbomole-format fixtures cut from the reference's own data/basis files, and (b) an analytic property: every renormalised
contraction has unit self-overlap."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN

SYMS = ["O", "H", "H", "C", "N", "S"]
XYZ = np.array([[0.0, 0.0, 0.2], [0.0, 1.4, -0.9], [0.0, -1.4, -0.9], [3.0, 0.1, 0.0], [-2.5, 0.3, 1.0], [0.5, 4.0, -1.0]])


def _dfact(n):
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


@pytest.mark.parametrize("spherical", [True, False])
def test_def2_svp_file_matches_the_python_producer(spherical):
    from serenity_b200.inputs.basis import build_shell_table, load_basis_set
    from serenity_b200.xc import shell_table_from_file
    tab, atom_of_bf = shell_table_from_file(os.path.join(GOLDEN, "basis_fixture_DEF2-SVP"), "def2-svp", SYMS[:5], XYZ[:5], spherical)
    assert set(load_basis_set("def2-svp")) >= {"h", "c", "n", "o"}
    ref = build_shell_table(SYMS[:5], XYZ[:5], "def2-svp", spherical=spherical)
    assert tab.nbf == ref.nbf and tab.nshell == ref.nshell
    for name in ("l", "pure", "nprim", "prim_off", "first_bf"):
        assert np.array_equal(getattr(tab, name), getattr(ref, name)), name
    assert np.array_equal(tab.centre, ref.centre) and np.array_equal(tab.alpha, ref.alpha)
    assert np.abs(tab.coeff - ref.coeff).max() <= 1e-14 * np.abs(ref.coeff).max()
    assert np.abs(tab.normfac - ref.normfac).max() <= 1e-15
    from serenity_b200.inputs.basis import atom_indices_of_basis
    assert np.array_equal(atom_of_bf, atom_indices_of_basis(ref, XYZ[:5]))


def test_unit_norm_and_fortran_exponents():
    """self-overlap of every contraction = 1 (libint2 renorm); sulfur (d shell, 5-fold contractions) and the 6-31G* file"""
    from serenity_b200.xc import shell_table_from_file
    for fixture, label, syms in (("basis_fixture_DEF2-SVP", "DEF2-SVP", ["S", "O"]), ("basis_fixture_6-31GS", "6-31GS", ["C", "O", "H"]),
                                 ("basis_fixture_DEF2-TZVP", "def2-TZVP", ["C", "H"])):
        tab, _ = shell_table_from_file(os.path.join(GOLDEN, fixture), label, syms, XYZ[:len(syms)], True)
        for s in range(tab.nshell):
            l, o, n = int(tab.l[s]), int(tab.prim_off[s]), int(tab.nprim[s])
            a, c = tab.alpha[o:o + n], tab.coeff[o:o + n]
            ovl = sum(c[p] * c[q] * _dfact(2 * l - 1) * math.pi ** 1.5 / (2.0 ** l * (a[p] + a[q]) ** (l + 1.5))
                      for p in range(n) for q in range(n))
            assert abs(ovl - 1.0) < 1e-13, (fixture, s)
    # 6-31G* carbon: s(6) s(3) s(1) p(3) p(1) d(1) in file order (sp shells are listed separately in the Turbomole format)
    tab, _ = shell_table_from_file(os.path.join(GOLDEN, "basis_fixture_6-31GS"), "6-31GS", ["C"], XYZ[:1], False)
    assert list(tab.l) == [0, 0, 0, 1, 1, 2] and list(tab.nprim) == [6, 3, 1, 3, 1, 1] and tab.nbf == 3 + 6 + 6
    assert np.allclose(tab.normfac[-6:], [1.0, math.sqrt(3), math.sqrt(3), 1.0, math.sqrt(3), 1.0])  # xx xy xz yy yz zz


def test_errors_are_worded_like_the_reference():
    from serenity_b200._lib import SerenityError
    from serenity_b200.xc import shell_table_from_file
    with pytest.raises(SerenityError, match="not defined for this element"):
        shell_table_from_file(os.path.join(GOLDEN, "basis_fixture_6-31GS"), "6-31GS", ["N"], XYZ[:1])
    with pytest.raises(SerenityError, match="Error while parsing basis file"):
        shell_table_from_file(os.path.join(GOLDEN, "no_such_file"), "6-31GS", ["H"], XYZ[:1])
    with pytest.raises(SerenityError, match="not defined for this element"):
        shell_table_from_file(os.path.join(GOLDEN, "basis_fixture_6-31GS"), "DEF2-SVP", ["H"], XYZ[:1])


def test_label_matching_is_anchored_where_the_reference_regex_is_not(tmp_path):
    """Deliberate differences from BasisFunctionProvider.cpp:61-96, documented in basis_provider.cpp: the element must start a line
    and the label must end at white space, so a label that is a prefix of another one (def2-SVP / def2-SVPD) and an element symbol
    that ends another word cannot resolve to the wrong entry (the reference's unanchored icase regex_search takes the first textual
    hit); any number of '*' / '#' lines may follow the header; lower-case Fortran exponents (d+01) are accepted as well."""
    from serenity_b200._lib import SerenityError
    from serenity_b200.xc import shell_table_from_file
    text = """$basis
*
h def2-SVPD
# h     (5s2p) / [3s2p]     {311/11}
*
    1  s
      9.0   1.0
    1  p
      0.9   1.0
*
h def2-SVP
# h     (4s1p) / [2s1p]     {31/1}
# a second comment line
*
    2  s
      0.13010701D+02      0.19682158d-01
      0.19622572D+01      0.13796524D+00
    1  s
      0.12179496D+00      1.0
*
bh def2-SVP
*
    1  s
      5.0   1.0
*
$end
"""
    path = tmp_path / "basis_prefix"
    path.write_text(text)
    tab, _ = shell_table_from_file(str(path), "def2-SVP", ["H"], XYZ[:1], True)
    assert list(tab.l) == [0, 0] and list(tab.nprim) == [2, 1]          # the def2-SVP entry, not the def2-SVPD one before it
    assert np.allclose(tab.alpha, [13.010701, 1.9622572, 0.12179496])    # D+02 and d-01 both read as exponents
    tab, _ = shell_table_from_file(str(path), "DEF2-SVPD", ["H"], XYZ[:1], True)
    assert list(tab.l) == [0, 1]
    with pytest.raises(SerenityError, match="not defined for this element"):
        shell_table_from_file(str(path), "def2-SV", ["H"], XYZ[:1], True)  # a prefix of a label is not a label


@pytest.mark.gpu
def test_gpu_build_from_basis_file():
    """geometry + basis file -> sxc_add_basis_from_table -> XC build == the build from the Python-made table"""
    import ctypes as C
    from serenity_b200 import _lib
    from serenity_b200.inputs import make_config
    from serenity_b200.inputs.configs import FUNCTIONALS
    from serenity_b200.xc import XCContext
    cfg = make_config("h2o", 2)
    sub = cfg.subsystems[0]
    ctx = XCContext(0)
    g = ctx.set_grid(cfg.xyz, cfg.w, 128)
    f = ctx.set_functional(*FUNCTIONALS["PBE"])
    V_ref, E_ref, _ = ctx.build_xc(g, ctx.add_basis(sub.basis, 1e-9), f, sub.P)
    lib = _lib.load()
    names = (C.c_char_p * len(sub.symbols))(*[s.encode() for s in sub.symbols])
    coords = np.ascontiguousarray(sub.coords, dtype=np.float64)
    h, b = C.c_void_p(), C.c_int(-1)
    assert lib.sxc_shell_table_from_file(os.path.join(GOLDEN, "basis_fixture_DEF2-SVP").encode(), b"DEF2-SVP", len(sub.symbols), names,
                                         coords.ctypes.data_as(C.c_void_p), 1, C.byref(h)) == 0
    assert lib.sxc_add_basis_from_table(ctx._h, h, 1e-9, C.byref(b)) == 0
    lib.sxc_shell_table_free(h)
    V, E, _ = ctx.build_xc(g, b.value, f, sub.P)
    assert np.abs(V - V_ref).max() <= 1e-13 and abs(E - E_ref) <= 1e-13
    ctx.close()


def test_parser_round_trip_on_generated_files(tmp_path):
    """Property test of the C++ parser (hypothesis): random element entries written in Turbomole format with varying white
    space, D / E exponents and optional comment lines come back with exactly the written exponents, contraction lengths and
    angular momenta, in file order, atom-major."""
    from hypothesis import given, settings, strategies as st
    from serenity_b200.xc import shell_table_from_file

    shell = st.tuples(st.sampled_from("spdfg"), st.lists(st.tuples(st.floats(1e-3, 1e5), st.floats(0.05, 2.0)),
                                                          min_size=1, max_size=6))
    entry = st.lists(shell, min_size=1, max_size=5)

    @settings(max_examples=40, deadline=None, derandomize=True)  # positive coefficients: no cancelling contractions
    @given(entries=st.lists(entry, min_size=1, max_size=3), fortran=st.booleans(), comment=st.booleans(), pad=st.integers(1, 6))
    def run(entries, fortran, comment, pad):
        names = ["h", "c", "o"][:len(entries)]
        lines = ["# generated", "$basis", "*"]
        for el, shells in zip(names, entries):
            lines.append("%s%sTEST-BASIS" % (el, " " * pad))
            if comment:
                lines.append("# %s  (generated)" % el)
            lines.append("*")
            for typ, prim in shells:
                lines.append("%s%d  %s" % (" " * pad, len(prim), typ))
                for a, c in prim:
                    fa, fc = "%.17e" % a, "%.17e" % c
                    if fortran:
                        fa, fc = fa.replace("e", "D"), fc.replace("e", "D")
                    lines.append("%s%s%s%s" % (" " * pad, fa, " " * pad, fc))
            lines.append("*")
        lines.append("$end")
        path = tmp_path / "gen_basis"
        path.write_text("\n".join(lines) + "\n")
        syms = [n.upper() for n in reversed(names)]          # atoms in another order than the file
        xyz = np.arange(3.0 * len(syms)).reshape(-1, 3)
        tab, atom_of_bf = shell_table_from_file(str(path), "test-basis", syms, xyz, True)
        want = [sh for n in reversed(names) for sh in entries[names.index(n)]]
        assert tab.nshell == len(want)
        assert list(tab.l) == ["spdfg".index(t) for t, _ in want]
        assert list(tab.nprim) == [len(p) for _, p in want]
        assert np.array_equal(tab.alpha, np.array([a for _, p in want for a, _ in p]))
        assert tab.nbf == sum(2 * "spdfg".index(t) + 1 for t, _ in want) == len(atom_of_bf)
        assert np.all(np.diff(atom_of_bf) >= 0) and atom_of_bf[-1] == len(syms) - 1
        # libint renormalisation: unit self-overlap of every contraction whose coefficients do not cancel
        for s in range(tab.nshell):
            l, o, n = int(tab.l[s]), int(tab.prim_off[s]), int(tab.nprim[s])
            a, c = tab.alpha[o:o + n], tab.coeff[o:o + n]
            if not np.all(np.isfinite(c)):
                continue
            ovl = sum(c[p] * c[q] * _dfact(2 * l - 1) * math.pi ** 1.5 / (2.0 ** l * (a[p] + a[q]) ** (l + 1.5))
                      for p in range(n) for q in range(n))
            assert abs(ovl - 1.0) < 1e-9

    run()
