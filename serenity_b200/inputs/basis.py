"""Shell tables: what Serenity's BasisController/Shell hand to the grid path.

Follows the reference's input producers (not part of the hot path):
  * src/basis/Shell.cpp:29-47          Shell = libint2::Shell + Cartesian norm factors
  * src/integrals/Normalization.h:67-87  finalNormalization(ax,ay,az)
  * libint2 2.7.0-beta.6 Shell::renorm  (un-vendored; algorithm restated)
  * src/basis/BasisController.cpp:62-68 extendedIndex(shell)
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "basis_data")


def _dfact(n: int) -> float:
    """(n)!! with (-1)!! = 1."""
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def renormalise(l: int, exps, coefs) -> np.ndarray:
    """libint2::Shell::renorm: primitive normalisation followed by unit normalisation of the contraction."""
    a = np.asarray(exps, dtype=np.float64)
    d = np.asarray(coefs, dtype=np.float64)
    df = _dfact(2 * l - 1)
    c = d * np.sqrt(2.0 ** l * (2.0 * a) ** (l + 1.5) / (math.pi ** 1.5 * df))
    n2 = 0.0
    for p in range(len(a)):
        for q in range(len(a)):
            n2 += c[p] * c[q] * df * math.pi ** 1.5 / (2.0 ** l * (a[p] + a[q]) ** (l + 1.5))
    return c / math.sqrt(n2)


def cartesian_norm_factors(l: int) -> np.ndarray:
    """sqrt((2l-1)!! / ((2a-1)!!(2b-1)!!(2c-1)!!)) in the order a = l..0, b = l-a..0 (Shell.cpp:37-47)."""
    out = []
    for a in range(l, -1, -1):
        for b in range(l - a, -1, -1):
            c = l - a - b
            out.append(math.sqrt(_dfact(2 * l - 1) / (_dfact(2 * a - 1) * _dfact(2 * b - 1) * _dfact(2 * c - 1))))
    return np.asarray(out)


def nfunc(l: int, pure: bool) -> int:
    return 2 * l + 1 if pure else (l + 1) * (l + 2) // 2


@dataclass
class ShellTable:
    """Flat shell table (structure of arrays) - the argument layout of sxc_add_basis."""
    l: np.ndarray         # int32 [nshell]
    pure: np.ndarray      # int32 [nshell]
    nprim: np.ndarray     # int32 [nshell]
    prim_off: np.ndarray  # int32 [nshell]
    first_bf: np.ndarray  # int32 [nshell]
    centre: np.ndarray    # float64 [nshell, 3] (bohr)
    alpha: np.ndarray     # float64 [sum nprim]
    coeff: np.ndarray     # float64 [sum nprim]  renormalised
    normfac: np.ndarray   # float64 [nbf]
    nbf: int

    @property
    def nshell(self) -> int:
        return int(self.l.shape[0])


def shell_table_from_list(shells) -> ShellTable:
    """shells: iterable of dicts {l, pure, exps, coefs (raw), centre (bohr)}."""
    l, pure, nprim, off, first, cen, al, co, nf = [], [], [], [], [], [], [], [], []
    nbf = 0
    npr = 0
    for sh in shells:
        ll, pp = int(sh["l"]), bool(sh.get("pure", True))
        l.append(ll)
        pure.append(1 if pp else 0)
        nprim.append(len(sh["exps"]))
        off.append(npr)
        first.append(nbf)
        cen.append(list(sh["centre"]))
        al.extend(sh["exps"])
        co.extend(renormalise(ll, sh["exps"], sh["coefs"]))
        nf.extend(np.ones(2 * ll + 1) if pp else cartesian_norm_factors(ll))
        nbf += nfunc(ll, pp)
        npr += len(sh["exps"])
    return ShellTable(np.asarray(l, np.int32), np.asarray(pure, np.int32), np.asarray(nprim, np.int32),
                      np.asarray(off, np.int32), np.asarray(first, np.int32),
                      np.ascontiguousarray(np.asarray(cen, np.float64).reshape(-1, 3)),
                      np.asarray(al, np.float64), np.asarray(co, np.float64), np.asarray(nf, np.float64), nbf)


_BASIS_CACHE: dict = {}


def load_basis_set(name: str) -> dict:
    key = name.lower()
    if key not in _BASIS_CACHE:
        with open(os.path.join(_DATA, key + ".json")) as f:
            _BASIS_CACHE[key] = json.load(f)["elements"]
    return _BASIS_CACHE[key]


def build_shell_table(symbols, coords_bohr, basis_name: str, spherical: bool = True) -> ShellTable:
    """Atom-major, file order of shells (src/basis/AtomCenteredBasisController; makeSphericalBasis default true)."""
    bs = load_basis_set(basis_name)
    shells = []
    for sym, xyz in zip(symbols, np.asarray(coords_bohr, dtype=np.float64)):
        for sh in bs[sym.lower()]:
            shells.append({"l": sh["l"], "pure": spherical, "exps": sh["exps"], "coefs": sh["coefs"],
                           "centre": xyz.tolist()})
    return shell_table_from_list(shells)


def atom_indices_of_basis(tab: ShellTable, coords_bohr) -> np.ndarray:
    """BasisController::getAtomIndicesOfBasis (used by FuncPotential::getGeomGradients, FuncPotential.cpp:127):
    index of the atom every basis function sits on (shell centre == atom position)."""
    coords = np.asarray(coords_bohr, dtype=np.float64)
    out = np.zeros(tab.nbf, dtype=np.int32)
    for sh in range(tab.nshell):
        atom = int(np.argmin(np.linalg.norm(coords - tab.centre[sh], axis=1)))
        n = nfunc(int(tab.l[sh]), bool(tab.pure[sh]))
        out[tab.first_bf[sh]: tab.first_bf[sh] + n] = atom
    return out
