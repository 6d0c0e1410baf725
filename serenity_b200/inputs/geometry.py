"""Synthetic molecular geometries of the BASELINE configs.

All generators are deterministic (fixed seeds) and return (symbols, coords_bohr).
xyz data in Angstrom; BOHR as src/parameters/Constants.h:63,79-81.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

BOHR_TO_ANGSTROM = 5.29177210544e-11 * 1.0e10
ANGSTROM_TO_BOHR = 1.0 / BOHR_TO_ANGSTROM

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "basis_data", "geometries.json")
_Z = {"H": 1, "C": 6, "N": 7, "O": 8}


def atomic_numbers(symbols):
    return [_Z[s.capitalize()] for s in symbols]


def n_electrons(symbols) -> int:
    return int(sum(atomic_numbers(symbols)))


def _stored(name):
    with open(_DATA) as f:
        atoms = json.load(f)["geometries"][name]
    syms = [a[0].capitalize() for a in atoms]
    xyz = np.asarray([a[1:4] for a in atoms], dtype=np.float64) * ANGSTROM_TO_BOHR
    return syms, xyz


def water():
    """cfg 1: data/xyzfiles/water.xyz."""
    return _stored("water")


def water_dimer():
    """cfg 4 (small case): data/xyzfiles/water_dimer.xyz; atoms 0-2 / 3-5 are the two subsystems."""
    return _stored("water_dimer")


def tetracene():
    """cfg 2: ideal planar tetracene C18H12, four linearly fused hexagons, r_CC = 1.40 A, r_CH = 1.09 A."""
    rcc, rch = 1.40, 1.09
    w = rcc * math.sqrt(3.0)  # ring width along the long axis
    carbons = []
    for ring in range(4):
        cx = ring * w
        for k in range(6):
            ang = math.radians(90.0 + 60.0 * k)  # vertices at the top and bottom, flat sides shared
            carbons.append((cx + rcc * math.cos(ang), rcc * math.sin(ang), 0.0))
    # remove duplicates of the shared edges
    uniq = []
    for c in carbons:
        if not any(math.dist(c, u) < 1e-6 for u in uniq):
            uniq.append(c)
    assert len(uniq) == 18, len(uniq)
    hyd = []
    for c in uniq:
        nb = [u for u in uniq if 1e-6 < math.dist(c, u) < rcc * 1.1]
        if len(nb) == 2:  # CH carbon: H points away from the two neighbours
            v = np.asarray(c) * 2 - np.asarray(nb[0]) - np.asarray(nb[1])
            v /= np.linalg.norm(v)
            hyd.append(tuple(np.asarray(c) + rch * v))
    assert len(hyd) == 12, len(hyd)
    syms = ["C"] * 18 + ["H"] * 12
    xyz = np.asarray(uniq + hyd, dtype=np.float64)
    xyz -= xyz.mean(axis=0)
    return syms, xyz * ANGSTROM_TO_BOHR


def _random_rotation(rng) -> np.ndarray:
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def water_cluster(n_side: int = 4, spacing_angstrom: float = 3.1, seed: int = 20240601):
    """cfg 3: n_side^3 copies of water.xyz on a cubic lattice, each rotated by a seeded random SO(3) element."""
    syms0, xyz0 = water()
    xyz0 = xyz0 - xyz0.mean(axis=0)
    rng = np.random.default_rng(seed)
    syms, xyz = [], []
    for i in range(n_side):
        for j in range(n_side):
            for k in range(n_side):
                R = _random_rotation(rng)
                shift = np.array([i, j, k], dtype=np.float64) * spacing_angstrom * ANGSTROM_TO_BOHR
                xyz.append(xyz0 @ R.T + shift)
                syms.extend(syms0)
    return syms, np.concatenate(xyz, axis=0)


def peptide_stand_in(n_copies: int = 8, seed: int = 20240601):
    """cfg 5: 8 translated copies of gly-gly-gly (24 atoms) + 8 H2O = 216 atoms, copies >= 4 A apart."""
    syms0, xyz0 = _stored("gly-gly-gly")
    wsyms, wxyz = water()
    wxyz = wxyz - wxyz.mean(axis=0)
    xyz0 = xyz0 - xyz0.mean(axis=0)
    extent = xyz0.max(axis=0) - xyz0.min(axis=0)
    pitch = extent + 4.0 * ANGSTROM_TO_BOHR
    rng = np.random.default_rng(seed)
    syms, xyz = [], []
    cells = [(i, j, k) for i in range(2) for j in range(2) for k in range(2)][:n_copies]
    for (i, j, k) in cells:
        shift = np.array([i, j, k], dtype=np.float64) * pitch
        xyz.append(xyz0 + shift)
        syms.extend(syms0)
        # one water per copy, placed in the gap half a pitch further along x/y/z diagonal
        R = _random_rotation(rng)
        xyz.append(wxyz @ R.T + shift + 0.5 * pitch)
        syms.extend(wsyms)
    return syms, np.concatenate(xyz, axis=0)
