#!/usr/bin/env python3
"""Generate the committed golden fixtures from the reference tree (run in the build container).

/root/reference is read-only and does NOT exist on the GPU box; the tests only read the
JSON files this script writes.  Nothing here copies reference *source*; it extracts
 (1) the numeric expectations of the reference's own unit tests (known-answer vectors),
 (2) values of the reference's hard-coded solid-harmonic formulas at seeded random points
     (the formulas are evaluated, their text is not kept),
 (3) public basis-set data (def2-SVP / def2-TZVP, Weigend & Ahlrichs 2005) for H, C, N, O and
     three small xyz geometries used to build the synthetic BASELINE configs.

Usage: python tests/golden/make_golden_from_reference.py [/root/reference]
"""
import json
import math
import os
import random
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def _read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


# ----------------------------------------------------------------------------------------------
# (1) EXPECT_NEAR tables
# ----------------------------------------------------------------------------------------------
_ENV = {"sqrt": math.sqrt}


def _expect_near(text):
    """Yield (expected_value, observed_expression) of every EXPECT_NEAR(a, b, tol)."""
    for m in re.finditer(r"EXPECT_NEAR\(\s*(.+?),\s*(.+?),\s*expectedPrecision\s*\)\s*;", text, re.S):
        yield m.group(1).strip(), m.group(2).strip()


def golden_basis_functions():
    """src/data/grid/BasisFunctionOnGridController_test.cpp:43-483 (TINY grid, SMALL_MIXED basis)."""
    text = _read("src/data/grid/BasisFunctionOnGridController_test.cpp")
    text = text[text.index("TestBasisFunctionValuesAndDerivatives"):text.index("class SphBFOnGridTest")]
    out = {}
    for a, b in _expect_near(text):
        m = re.fullmatch(r"(values|derivatives\.(?:x|y|z)|hessian\.(?:xx|xy|xz|yy|yz|zz))\((\d+),\s*(\d+)\)(?:\s*/\s*sqrt\([^)]*\))?", b)  # one line divides the observed side (expected 0)
        if not m:
            continue
        key = m.group(1).replace("derivatives.", "d").replace("hessian.", "h")
        out.setdefault(key, []).append([int(m.group(2)), int(m.group(3)), float(eval(a, _ENV))])
    return {"source": "src/data/grid/BasisFunctionOnGridController_test.cpp:43-483", "tolerance": 1e-8,
            "block_size": 10, "radial_threshold": 1e-11, "entries": out}


def golden_density():
    """src/data/grid/DensityOnGridCalculator_test.cpp:43-254."""
    text = _read("src/data/grid/DensityOnGridCalculator_test.cpp")
    P = [[0.0] * 10 for _ in range(10)]
    for m in re.finditer(r"densityMatrix\((\d+),\s*(\d+)\)\s*=\s*([^;]+);", text):
        P[int(m.group(1))][int(m.group(2))] = float(eval(m.group(3), _ENV))
    out = {}
    for a, b in _expect_near(text):
        m = re.fullmatch(r"(densOnGrid|gradOnGrid\.(?:x|y|z)|hessOnGrid\.(?:xx|xy|xz|yy|yz|zz))\[(\d+)\]", b)
        key = m.group(1).replace("densOnGrid", "rho").replace("gradOnGrid.", "d").replace("hessOnGrid.", "h")
        out.setdefault(key, {})[int(m.group(2))] = float(eval(a, _ENV))
    out = {k: [v[i] for i in sorted(v)] for k, v in out.items()}
    return {"source": "src/data/grid/DensityOnGridCalculator_test.cpp:43-254", "tolerance": 5e-8,
            "block_size": 3, "radial_threshold": 0.0, "P": P, "expected": out}


def golden_scatter():
    """src/data/grid/ScalarOperatorToMatrixAdder_test.cpp:41-148."""
    text = _read("src/data/grid/ScalarOperatorToMatrixAdder_test.cpp")
    pot = {}
    for m in re.finditer(r"\b(pot|gradPot\.x|gradPot\.y|gradPot\.z)\[(\d+)\]\s*=\s*([^;]+);", text):
        pot.setdefault(m.group(1).replace("gradPot.", "g"), {})[int(m.group(2))] = float(eval(m.group(3), _ENV))
    pot = {k: [v[i] for i in sorted(v)] for k, v in pot.items()}
    entries = []
    for m in re.finditer(r"EXPECT_NEAR\(result\((\d+),\s*(\d+)\),\s*([-+0-9.eE]+),\s*expectedPrecision\)", text):
        entries.append([int(m.group(1)), int(m.group(2)), float(m.group(3))])
    return {"source": "src/data/grid/ScalarOperatorToMatrixAdder_test.cpp:41-148", "tolerance": 1e-8,
            "block_size": 3, "radial_threshold": 1e-11, "block_ave_threshold": 1e-11, "potential": pot,
            "entries": entries}


def golden_fixtures():
    """Grids and bases of src/testsupply/{GridController,BasisController}__TEST_SUPPLY.cpp."""
    return {
        "source": "src/testsupply/GridController__TEST_SUPPLY.cpp:31-46, BasisController__TEST_SUPPLY.cpp:36-80",
        "grids": {
            # Eigen '<<' fills row-major: rows are x, y, z
            "TINY": {"x": [0.0, 1.0, -2.0, 3.0], "y": [0.0, 0.0, 1.0, 0.5], "z": [0.0, 0.0, 1.0, -1.0],
                     "w": [1.0, 1.0, 1.0, 1.0]},
            "VERY_SMALL": {"x": [-1.0, 0.0, 0.0, 0.8, 0.1], "y": [0.0, 0.5, -0.4, 0.1, -0.1],
                           "z": [0.0, 0.4, -0.5, -0.1, 0.1], "w": [0.1, 0.2, 0.15, 0.15, 0.4]},
        },
        "bases": {
            "MINIMAL": [
                {"l": 0, "pure": False, "exps": [2.0], "coefs": [0.5], "centre": [0.0, 0.0, -1.0]},
                {"l": 0, "pure": False, "exps": [1.0], "coefs": [0.8], "centre": [0.0, 0.0, 1.0]},
            ],
            "SMALL_MIXED": [
                {"l": 0, "pure": False, "exps": [1.0, 2.0], "coefs": [1.0, 0.1], "centre": [0.0, 0.0, 0.0]},
                {"l": 1, "pure": False, "exps": [1.0, 2.0], "coefs": [1.0, 0.1], "centre": [1.0, 0.0, 0.0]},
                {"l": 2, "pure": False, "exps": [0.5, 1.2], "coefs": [0.8, 0.4], "centre": [-1.0, 2.5, 1.0]},
            ],
        },
    }


# ----------------------------------------------------------------------------------------------
# (2) harmonics of the reference evaluated at random points
# ----------------------------------------------------------------------------------------------
def golden_harmonics():
    src = _read("src/data/grid/BasisFunctionOnGridController.cpp")
    body = src[src.index("switch (angularMomentumOfMu) {"):src.index("Angular momentum too high")]
    cases = re.split(r"\bcase (\d+):", body)[1:]
    rng = random.Random(20240601)
    pts = [[rng.uniform(-1.5, 1.5) for _ in range(3)] for _ in range(6)]
    names = ["Y", "dYdx", "dYdy", "dYdz", "d2Ydxdx", "d2Ydxdy", "d2Ydxdz", "d2Ydydy", "d2Ydydz", "d2Ydzdz"]
    result = {"source": "src/data/grid/BasisFunctionOnGridController.cpp:441-1066 evaluated numerically",
              "points": pts, "names": names, "values": {}}
    for l_str, code in zip(cases[0::2], cases[1::2]):
        l = int(l_str)
        stmts = re.findall(r"\b(d2Yd[xyz]d[xyz]|dYd[xyz]|Y)\[(\d+)\]\s*=\s*([^;]+);", code)
        per_point = []
        for p in pts:
            env = {"sqrt": math.sqrt,
                   "x": [p[0] ** k for k in range(l + 3)], "y": [p[1] ** k for k in range(l + 3)],
                   "z": [p[2] ** k for k in range(l + 3)]}
            vals = {n: [0.0] * (2 * l + 1) for n in names}
            for name, idx, expr in stmts:
                vals[name][int(idx)] = float(eval(" ".join(expr.split()), env))
            per_point.append([vals[n] for n in names])
        result["values"][str(l)] = per_point
    return result


# ----------------------------------------------------------------------------------------------
# (3) basis-set data and geometries (public data, needed on the GPU box for the synthetic configs)
# ----------------------------------------------------------------------------------------------
_LNUM = {"s": 0, "p": 1, "d": 2, "f": 3, "g": 4, "h": 5, "i": 6}


def parse_turbomole_basis(rel, elements):
    """Turbomole-format parser (format as read by src/basis/BasisFunctionProvider.cpp)."""
    lines = _read(rel).splitlines()
    out = {}
    i = 0
    while i < len(lines):
        ln = lines[i].strip()
        m = re.fullmatch(r"([a-z]{1,2})\s+\S+", ln)
        if m and i > 0 and lines[i - 1].strip() == "*" and lines[i + 1].strip() == "*":
            el = m.group(1)
            i += 2
            shells = []
            while i < len(lines) and lines[i].strip() != "*":
                hm = re.fullmatch(r"(\d+)\s+([spdfghi])", lines[i].strip())
                if hm is None:  # blank/comment lines or an element block in another layout: skip the line
                    i += 1
                    continue
                nprim, l = int(hm.group(1)), _LNUM[hm.group(2)]
                exps, coefs = [], []
                for k in range(nprim):
                    a, c = lines[i + 1 + k].split()[:2]
                    exps.append(float(a.replace("D", "E")))
                    coefs.append(float(c.replace("D", "E")))
                shells.append({"l": l, "exps": exps, "coefs": coefs})
                i += 1 + nprim
            if el in elements:
                out[el] = shells
        i += 1
    return out


def geometries():
    geo = {}
    for name in ["water", "water_dimer", "gly-gly-gly"]:
        lines = _read(f"data/xyzfiles/{name}.xyz").splitlines()
        n = int(lines[0].split()[0])
        atoms = []
        for ln in lines[2:2 + n]:
            s, x, y, z = ln.split()[:4]
            atoms.append([s, float(x), float(y), float(z)])
        geo[name] = atoms
    return geo


def main():
    def dump(obj, path):
        with open(path, "w") as f:
            json.dump(obj, f, indent=1)
        print("wrote", os.path.relpath(path, REPO))

    dump(golden_fixtures(), os.path.join(HERE, "fixtures_testsupply.json"))
    dump(golden_basis_functions(), os.path.join(HERE, "basis_functions_ref.json"))
    dump(golden_density(), os.path.join(HERE, "density_ref.json"))
    dump(golden_scatter(), os.path.join(HERE, "scatter_ref.json"))
    dump(golden_harmonics(), os.path.join(HERE, "harmonics_ref.json"))
    data_dir = os.path.join(REPO, "serenity_b200", "inputs", "basis_data")
    els = {"h", "c", "n", "o"}
    dump({"name": "def2-SVP", "source": "data/basis/DEF2-SVP", "elements": parse_turbomole_basis("data/basis/DEF2-SVP", els)},
         os.path.join(data_dir, "def2-svp.json"))
    dump({"name": "def2-TZVP", "source": "data/basis/DEF2-TZVP", "elements": parse_turbomole_basis("data/basis/DEF2-TZVP", els)},
         os.path.join(data_dir, "def2-tzvp.json"))
    dump({"source": "data/xyzfiles/*.xyz (Angstrom)", "geometries": geometries()},
         os.path.join(data_dir, "geometries.json"))


if __name__ == "__main__":
    main()
