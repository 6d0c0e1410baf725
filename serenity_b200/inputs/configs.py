"""The five BASELINE.json configs as concrete synthetic inputs."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import geometry as geo
from .basis import ShellTable, build_shell_table
from .density import random_density_matrix
from .grid import molecular_grid

# BASIC_FUNCTIONALS enum values, src/dft/functionals/BasicFunctionals.h:39-...
X_SLATER, C_VWN, K_TF, X_B88, X_B88_CORR, X_PBE, C_LYP, C_P86, C_PBE, K_PW91, K_LLP = 2, 45, 66, 80, 81, 135, 184, 193, 197, 283, 286

# composite definitions, src/dft/functionals/CompositeFunctionals.cpp:224-340 (XCFun route)
FUNCTIONALS = {
    "LDA": ([X_SLATER, C_VWN], [1.0, 1.0]),
    "PBE": ([X_PBE, C_PBE], [1.0, 1.0]),
    "BLYP": ([X_B88, C_LYP], [1.0, 1.0]),
    "BP86": ([X_B88, C_P86], [1.0, 1.0]),
    "B3LYP": ([X_SLATER, X_B88_CORR, C_LYP, C_VWN], [0.80, 0.72, 0.81, 0.19]),  # 20 % exact exchange is not grid work
    "PBE0": ([X_PBE, C_PBE], [0.75, 1.0]),
    "SLATER": ([X_SLATER], [1.0]),
    "TF": ([K_TF], [1.0]),
    "PW91K": ([K_PW91], [1.0]),
    "LLP91K": ([K_LLP], [1.0]),
}


@dataclass
class Subsystem:
    symbols: list
    coords: np.ndarray
    basis: ShellTable
    P: np.ndarray
    n_electrons: int


@dataclass
class Config:
    name: str
    description: str
    functional: str
    xyz: np.ndarray          # [N,3] grid points (bohr)
    w: np.ndarray            # [N]
    subsystems: list = field(default_factory=list)  # 1 entry for KS-DFT, >= 2 for FDE (0 = active)
    nadd_kin: str | None = None  # second NAdd functional for FDE

    @property
    def npts(self) -> int:
        return int(self.w.shape[0])


def _subsystem(symbols, coords, basis_name, seed):
    tab = build_shell_table(symbols, coords, basis_name, spherical=True)
    nel = geo.n_electrons(symbols)
    return Subsystem(list(symbols), np.asarray(coords), tab, random_density_matrix(tab.nbf, nel, seed), nel)


def geometry_of(name: str):
    """(symbols, coordinates in bohr) of a named synthetic system."""
    n = name.lower()
    if n == "h2o":
        return geo.water()
    if n == "tetracene":
        return geo.tetracene()
    if n in ("water64", "water8", "water27"):
        return geo.water_cluster({"water64": 4, "water27": 3, "water8": 2}[n])
    if n == "peptide":
        return geo.peptide_stand_in()
    raise ValueError("unknown geometry " + name)


def make_config(name: str, acc: int | None = None, grid=None) -> Config:
    """name in {h2o, tetracene, water64, fde_dimer, fde_water64, peptide} (+ small test variants).
    grid = (xyz, w): reuse an already generated grid of the same geometry and accuracy (water64 and fde_water64 share one)."""
    n = name.lower()
    mg = molecular_grid if grid is None else (lambda *_a, **_k: grid)
    if n == "h2o":
        s, c = geo.water()
        xyz, w = mg(s, c, acc or 4)
        return Config(n, "H2O PBE/def2-SVP, grid accuracy %d" % (acc or 4), "PBE", xyz, w, [_subsystem(s, c, "def2-svp", 7)])
    if n == "tetracene":
        s, c = geo.tetracene()
        xyz, w = mg(s, c, acc or 6)
        return Config(n, "tetracene C18H12 B3LYP/def2-TZVP, grid accuracy %d" % (acc or 6), "B3LYP", xyz, w,
                      [_subsystem(s, c, "def2-tzvp", 7)])
    if n in ("water64", "water8", "water27"):
        side = {"water64": 4, "water27": 3, "water8": 2}[n]
        s, c = geo.water_cluster(side)
        xyz, w = mg(s, c, acc or 4)
        return Config(n, "(H2O)%d PBE/def2-SVP, grid accuracy %d" % (side ** 3, acc or 4), "PBE", xyz, w,
                      [_subsystem(s, c, "def2-svp", 7)])
    if n == "fde_dimer":
        s, c = geo.water_dimer()
        xyz, w = mg(s, c, acc or 4)
        return Config(n, "freeze-and-thaw FDE water dimer, PW91k + PBE, supersystem grid accuracy %d" % (acc or 4), "PBE",
                      xyz, w, [_subsystem(s[:3], c[:3], "def2-svp", 7), _subsystem(s[3:], c[3:], "def2-svp", 8)], "PW91K")
    if n in ("fde_water64", "fde_water8"):
        side = 4 if n == "fde_water64" else 2
        s, c = geo.water_cluster(side)
        xyz, w = mg(s, c, acc or 4)
        h = len(s) // 2
        return Config(n, "freeze-and-thaw FDE (H2O)%d split in two halves, PW91k + PBE" % (side ** 3), "PBE", xyz, w,
                      [_subsystem(s[:h], c[:h], "def2-svp", 7), _subsystem(s[h:], c[h:], "def2-svp", 8)], "PW91K")
    if n == "peptide":
        s, c = geo.peptide_stand_in()
        xyz, w = mg(s, c, acc or 6)
        return Config(n, "216-atom peptide stand-in PBE/def2-SVP, grid accuracy %d" % (acc or 6), "PBE", xyz, w,
                      [_subsystem(s, c, "def2-svp", 7)])
    raise ValueError("unknown config " + name)
