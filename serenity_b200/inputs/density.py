"""Synthetic closed-shell density matrices.

P = 2 C C^T with C (nb x n_occ) Gaussian random (seed 7): symmetric, positive semi-definite, so rho >= 0
everywhere; callers rescale P by N_el / integral(rho) once a density integral is available
(rho is linear in P).  RESTRICTED P is the total density matrix (occupation 2).
"""
from __future__ import annotations

import numpy as np


def random_density_matrix(nbf: int, n_electrons: int, seed: int = 7) -> np.ndarray:
    nocc = max(1, n_electrons // 2)
    rng = np.random.default_rng(seed)
    C = rng.normal(size=(nbf, nocc)) / np.sqrt(nbf)
    P = 2.0 * (C @ C.T)
    return np.asfortranarray(0.5 * (P + P.T))
