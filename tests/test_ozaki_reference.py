"""INT8-slice reference of the contractions (oracle/ozaki.c) - the fixed slicing convention the planned tcgen05 kind::i8
kernels will be compared with bit-exactly.  Here: its own invariants."""
import numpy as np


def test_slices_reconstruct_rows_and_products_are_exact_integers():
    from oracle import pyoracle as orc
    rng = np.random.default_rng(0)
    X = rng.standard_normal((37, 128)) * np.exp(rng.uniform(-12, 3, (37, 1)))
    X[5] = 0.0                      # an all-zero row (padding rows of a tile)
    X[6, :] = 0.0
    X[6, 17] = 0.25                 # a power of two: |x| 2^-e = 1 exactly
    for k in (3, 5, 7):
        S, e = orc.ozaki_slice_rows(X, k)
        assert S.dtype == np.int8 and np.abs(S.astype(int)).max() <= 127
        rec = sum(S[i].astype(np.float64) * 128.0 ** -(i + 1) for i in range(k)) * (2.0 ** e)[:, None]
        bound = (2.0 ** e)[:, None] * 128.0 ** -k          # clamp at +-127 costs at most one unit of the last slice
        assert np.all(np.abs(rec - X) <= bound)
        assert e[5] == 0 and not S[:, 5].any() and e[6] == -2
    # exact integer accumulators: compare with int64 numpy on the same slices
    A = rng.standard_normal((24, 96))
    B = rng.standard_normal((40, 96))
    k = 5
    C, acc = orc.ozaki_matmul(A, B, k)
    Sa, ea = orc.ozaki_slice_rows(A, k)
    Sb, eb = orc.ozaki_slice_rows(B, k)
    for d in range(k):
        ref = sum(Sa[i].astype(np.int64) @ Sb[d - i].astype(np.int64).T for i in range(d + 1))
        assert np.array_equal(acc[d].astype(np.int64), ref)
    err = np.abs(C - A @ B.T).max()
    assert err <= 96 * (k + 1) * 128.0 ** -k * 4.0 ** 2 and err > 0.0   # 2^(eA + eB) <= 16 for unit normals here


def test_sliced_density_contraction_meets_the_tolerances_with_five_slices():
    """one block of the H2O grid: B = phi_s P_s with 5 slices reproduces rho to 1e-9 relative (tools/ozaki_study.py runs the
    whole build this way)"""
    from oracle import pyoracle as orc
    from serenity_b200.inputs import make_config
    cfg = make_config("h2o", 2)
    sub = cfg.subsystems[0]
    ob, og = orc.Basis(sub.basis), orc.Grid(cfg.xyz, cfg.w, 128)
    (val, dx, dy, dz), neg, _ = orc.basis_block(ob, og, 1e-9, 1, 3)
    sig = np.nonzero(neg == 0)[0]
    f = np.ascontiguousarray(val[:, sig])               # [points, functions]
    Ps = np.ascontiguousarray(sub.P[np.ix_(sig, sig)])
    # operands K-contiguous: rows of phi^T are functions over points ... the contraction index here is the function
    Bm, _ = orc.ozaki_matmul(f, Ps.T.copy(), 5)        # B[p, j] = sum_i f[p, i] P[i, j]
    rho = (Bm * f).sum(axis=1)
    ref = ((f @ Ps) * f).sum(axis=1)
    assert np.abs(rho - ref).max() <= 1e-9 * np.abs(ref).max()
