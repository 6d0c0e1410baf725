#!/usr/bin/env python
"""Writes small Turbomole-format basis-set fixtures (tests/golden/basis_fixture_<LABEL>) for the C++ front end
(serenity_b200/csrc/basis_provider.cpp).  The published basis-set parameters (Weigend & Ahlrichs 2005; Hehre, Ditchfield,
Pople 1972) of a few light elements are READ from the reference's data/basis/<LABEL> files and re-emitted by this script's own
writer - one file in plain decimal notation, DEF2-TZVP with Fortran D exponents (the D+/D- -> E+/E- replacement of
BasisFunctionProvider.cpp:93 is otherwise never exercised) and with the optional "# element (..) / [..]" comment line.
Run in the build container (needs /root/reference); the fixtures are committed."""
import os
import re
import sys

REF = "/root/reference/data/basis"
OUT = os.path.dirname(os.path.abspath(__file__))
WANT = {"DEF2-SVP": ["h", "c", "n", "o", "s"], "6-31GS": ["h", "c", "o"], "DEF2-TZVP": ["h", "c"]}


def parse(text, el, label):
    m = re.search(r"^%s\s+%s\s*$" % (el, re.escape(label)), text, re.I | re.M)
    if not m:
        sys.exit("no entry for %s in %s" % (el, label))
    star1 = text.index("*", m.end())
    star2 = text.index("*", star1 + 1)
    tok = [t for ln in text[star1 + 1:star2].splitlines() if not ln.strip().startswith("#") for t in ln.split()]
    shells, i = [], 0
    while i < len(tok):
        n, typ = int(tok[i]), tok[i + 1]
        prim = [(float(tok[i + 2 + 2 * k].replace("D", "E").replace("d", "e")),
                 float(tok[i + 3 + 2 * k].replace("D", "E").replace("d", "e"))) for k in range(n)]
        shells.append((typ, prim))
        i += 2 + 2 * n
    return shells


def fmt(x, fortran):
    s = "%.16e" % x
    return s.replace("e", "D") if fortran else repr(x)


def main():
    for label, elements in WANT.items():
        text = open(os.path.join(REF, label)).read()
        fortran = label == "DEF2-TZVP"
        out = ["# %s fixture written by tests/golden/make_basis_fixture.py (parameters read from the reference's data/basis/%s)"
               % (label, label), "$basis", "*"]
        for el in elements:
            shells = parse(text, el, label)
            out.append("%s   %s" % (el, label))
            if fortran:
                out.append("# %s     (%s)" % (el, "".join("%d%s" % (sum(1 for t, _ in shells if t == c), c) for c in "spdfg"
                                                          if any(t == c for t, _ in shells))))
            out.append("*")
            for typ, prim in shells:
                out.append("%5d  %s" % (len(prim), typ))
                for a, c in prim:
                    out.append("  %s   %s" % (fmt(a, fortran), fmt(c, fortran)))
            out.append("*")
        out.append("$end")
        with open(os.path.join(OUT, "basis_fixture_" + label), "w") as f:
            f.write("\n".join(out) + "\n")
        print("wrote basis_fixture_" + label)


if __name__ == "__main__":
    main()
