#!/usr/bin/env python
"""Cuts a small Turbomole-format fixture out of the reference's basis-set files (data/basis/DEF2-SVP, 6-31GS, DEF2-TZVP):
the header + the entries of a few light elements each, written to tests/golden/basis_fixture_<LABEL>.  The files carry
published basis-set parameters (Weigend & Ahlrichs 2005; Hehre, Ditchfield, Pople 1972 ...), i.e. data, no code.
Run in the build container (needs /root/reference); the fixtures are committed."""
import os
import re
import sys

REF = "/root/reference/data/basis"
OUT = os.path.dirname(os.path.abspath(__file__))
WANT = {"DEF2-SVP": ["h", "c", "n", "o", "s"], "6-31GS": ["h", "c", "o"], "DEF2-TZVP": ["h", "c"]}


def main():
    for label, elements in WANT.items():
        text = open(os.path.join(REF, label)).read()
        out = ["# fixture cut from the reference's data/basis/%s by tests/golden/make_basis_fixture.py" % label, "$basis", "*"]
        for el in elements:
            m = re.search(r"^%s\s+%s\s*$" % (el, re.escape(label)), text, re.I | re.M)
            if not m:
                sys.exit("no entry for %s in %s" % (el, label))
            star1 = text.index("*", m.end())
            star2 = text.index("*", star1 + 1)
            out.append(text[m.start():star2].rstrip("\n"))
            out.append("*")
        out.append("$end")
        with open(os.path.join(OUT, "basis_fixture_" + label), "w") as f:
            f.write("\n".join(out) + "\n")
        print("wrote basis_fixture_" + label)


if __name__ == "__main__":
    main()
