#!/usr/bin/env python
"""Feasibility study (CPU, numpy) for the step beyond the FP64 pipe: the two contractions of the XC build emulated with
INT8 slices (Ozaki splitting) as they would run on tcgen05 kind::i8 tiles with INT32 accumulation in TMEM.

Every operand row (phi, grad phi, G: one row = one basis function over the 128 points of a block; P_s: one column) is scaled
by a power of two to |x| < 1 and cut into k signed 7-bit slices x = sum_i x_i 2^(-7 (i + 1)); a product of slices is an exact
integer GEMM (K <= 512: |sum| < 2^23), and C ~ sum_{i + j < k} 2^(-7 (i + j + 2)) A_i B_j needs k (k + 1) / 2 of them.  The
script evaluates, block by block with the oracle's basis functions, what the north_star tolerances (1e-9 Eh in E_xc, 1e-8
max-abs in V_xc) demand:

  python tools/ozaki_study.py [h2o|water8|tetracene] [accuracy] [max blocks]

and prints one JSON line per slice count k: max relative error of rho, |dE_xc|, max |dV_xc| and the number of INT8 GEMMs per
FP64 GEMM.  Nothing here is product code; the exact products are done in float64 on integer-valued matrices."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as orc  # noqa: E402
from serenity_b200.inputs import make_config  # noqa: E402
from serenity_b200.inputs.configs import FUNCTIONALS  # noqa: E402


def slices(X, axis, k):
    """X scaled along `axis` (one power of two per row / column) and cut into k int8 slices; returns (list of slices, scale)."""
    amax = np.abs(X).max(axis=axis, keepdims=True)
    e = np.where(amax > 0, np.ceil(np.log2(np.where(amax > 0, amax, 1.0)) + 1e-12), 0.0)
    scale = 2.0 ** e                      # |X / scale| <= 1
    R = X / scale
    out = []
    for _ in range(k):
        R = R * 128.0
        S = np.round(R)                   # integers in [-128, 128]; the remainder is in [-1/2, 1/2]
        S = np.clip(S, -127, 127)
        out.append(S)
        R = R - S
    return out, scale


def sliced_matmul(A, B, k):
    """A [m, K] (rows scaled), B [K, n] (columns scaled): sum over slice pairs with i + j < k of exact integer products."""
    As, sa = slices(A, 1, k)
    Bs, sb = slices(B, 0, k)
    C = np.zeros((A.shape[0], B.shape[1]))
    for i in range(k):
        for j in range(k - i):
            C += (As[i] @ Bs[j]) * 2.0 ** (-7 * (i + j + 2))
    return C * sa * sb


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "h2o"
    acc = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    max_blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
    cfg = make_config(name, acc)
    sub = cfg.subsystems[0]
    nblk = min((cfg.npts + 127) // 128, max_blocks)
    npts = min(nblk * 128, cfg.npts)
    ob, og = orc.Basis(sub.basis), orc.Grid(cfg.xyz[:npts], cfg.w[:npts], 128)
    func = orc.Functional(*FUNCTIONALS[cfg.functional])
    nb = ob.nbf
    P = np.asarray(sub.P)
    V_ref, E_ref, _, _ = orc.build_xc(ob, og, func, P)
    rho_ref, g_ref, _, _ = orc.density_on_grid(ob, og, 1e-9, P, 1)
    blocks = []
    for b in range(nblk):
        (val, dx, dy, dz), neg, _ = orc.basis_block(ob, og, 1e-9, 1, b)
        sig = np.nonzero(neg == 0)[0]
        blocks.append((sig, val[:, sig], dx[:, sig], dy[:, sig], dz[:, sig]))
    for k in range(3, 9):
        rho = np.zeros(npts)
        grad = [np.zeros(npts) for _ in range(3)]
        for b, (sig, f, fx, fy, fz) in enumerate(blocks):
            if len(sig) == 0:
                continue
            lo = b * 128
            n = f.shape[0]
            # density: B = phi_s P_s (points x functions); the operand layout of the kernels is function-major, the scaling
            # unit is the function (a row of the tile, a column of P_s)
            Bm = sliced_matmul(f, P[np.ix_(sig, sig)], k) if k < 8 else f @ P[np.ix_(sig, sig)]
            rho[lo:lo + n] = (Bm * f).sum(axis=1)
            for c, d in enumerate((fx, fy, fz)):
                grad[c][lo:lo + n] = 2.0 * (Bm * d).sum(axis=1)
        E, out = orc.functional_on_grid(func, og.w, rho, *(grad if func.is_gga else (None,) * 3))
        V = np.zeros((nb, nb))
        for b, (sig, f, fx, fy, fz) in enumerate(blocks):
            if len(sig) == 0:
                continue
            lo = b * 128
            n = f.shape[0]
            w = og.w[lo:lo + n]
            a = w * out[1][lo:lo + n]
            G = 0.5 * a[:, None] * f
            if func.is_gga:
                for c, d in enumerate((fx, fy, fz)):
                    G += (w * out[2 + c][lo:lo + n])[:, None] * d
            T = sliced_matmul(f.T.copy(), G, k) if k < 8 else f.T @ G   # phi^T G: rows = functions, columns = functions
            V[np.ix_(sig, sig)] += T + T.T
        line = {"workload": cfg.description, "blocks": nblk, "slices": k if k < 8 else "fp64 (numpy, block by block)",
                "int8_gemms_per_fp64_gemm": k * (k + 1) // 2 if k < 8 else 0,
                "max_rel_drho": float(np.abs(rho - rho_ref).max() / np.abs(rho_ref).max()),
                "dE_xc": float(abs(E - E_ref)), "max_dV_xc": float(np.abs(V - V_ref).max())}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
