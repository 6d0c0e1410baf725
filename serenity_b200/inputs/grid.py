"""Atom-centred integration grids (input producer of the hot path).

Restates the construction of the reference so that the synthetic configs have the reference's point counts,
weights and block structure:
  * src/grid/construction/AtomGridFactory.cpp:78-198   radial counts, Ahlrichs M3 radial map (:236-255),
    pruned Lebedev shells (zones :170-190); Lebedev rules come from scipy.integrate.lebedev_rule
    (the reference ships Burkardt's table, sphere_lebedev_rule.cpp - same rules, different point order)
  * src/grid/construction/GridFactory.cpp:52-321        SSF / Becke partition weights, weight cut 1e-14
  * src/grid/HilbertRTreeSorting.cpp:46-214             locality sort: `hilbert_rtree_order` restates the reference's curve and
    order exactly (sort="reference"; pinned by its 2x2x2 test); the default of the synthetic configs is a plain 3-D Hilbert
    index with 10 bits per axis (same purpose, different orientation)
Blocks are runs of `blocksize` consecutive points of the returned order.
"""
from __future__ import annotations

import ctypes
import math
import os

import numpy as np
from scipy.integrate import lebedev_rule

_HERE = os.path.dirname(os.path.abspath(__file__))

# AtomGridFactory.cpp:39-52 (H..Ne shown; enough for the BASELINE configs) and :54-63
_AHLRICHS_ALPHA = {1: 0.8, 2: 0.9, 3: 1.8, 4: 1.4, 5: 1.3, 6: 1.1, 7: 0.9, 8: 0.9, 9: 0.9, 10: 0.9}
_CLEMENTI = {1: 1.00, 2: 0.59, 3: 3.16, 4: 2.12, 5: 1.64, 6: 1.27, 7: 1.06, 8: 0.91, 9: 0.79, 10: 0.72}
# Bragg-Slater radii in Angstrom (Slater 1964) for the BECKE flavour's size adjustment
_BRAGG_SLATER = {1: 0.25, 6: 0.70, 7: 0.65, 8: 0.60}
_RAD_ACC = [13, 13, 13, 14, 15, 16, 17]
# index -> Lebedev degree; point counts 6,14,26,38,50,74,86,110,146,170,194,230,266,302,350,434,590,770,...
_LEB_DEGREE = [3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 35, 41, 47, 53, 59, 65]
_LDVAL = [[4, 4, 4, 4, 4], [4, 4, 4, 7, 4], [4, 4, 7, 10, 7], [4, 7, 10, 13, 10], [7, 10, 13, 15, 13],
          [10, 13, 15, 16, 15], [13, 15, 16, 17, 16]]
_RANGES = [[0.25, 0.5, 1.0, 4.5], [0.1667, 0.5, 0.9, 3.5], [0.1, 0.4, 0.8, 2.5]]

_LEB_CACHE: dict = {}


def _lebedev(index: int):
    if index not in _LEB_CACHE:
        x, w = lebedev_rule(_LEB_DEGREE[index])
        _LEB_CACHE[index] = (np.ascontiguousarray(x.T), w / (4.0 * math.pi))  # weights normalised to 1
    return _LEB_CACHE[index]


def _row(z: int) -> int:
    return 1 if z <= 2 else 2 if z <= 10 else 3


def ahlrichs_radial(alpha: float, n: int):
    """AtomGridFactory.cpp:236-255 (M3 map with exponent 0.6, Chebyshev 2nd kind)."""
    i = np.arange(1, n + 1, dtype=np.float64)
    tmp = alpha / math.log(2.0)
    xi = np.cos(i * math.pi / (n + 1.0))
    ri = tmp * (xi + 1.0) ** 0.6 * np.log(2.0 / (1.0 - xi))
    ln = np.log((1.0 - xi) / 2.0)
    sq = np.sqrt((1.0 + xi) / (1.0 - xi))
    wi = (math.pi / (n + 1.0)) * (1.0 + xi) ** 1.8 * tmp ** 3 * (sq * ln * ln - 0.6 * ln ** 3 / sq)
    return ri[::-1].copy(), wi[::-1].copy()  # radPoints[nRadial - i]


def becke_radial(alpha: float, n: int):
    """AtomGridFactory.cpp:203-212."""
    i = np.arange(1, n + 1, dtype=np.float64)
    xi = np.cos(i * math.pi / (n + 1))
    w = np.sqrt((1.0 + xi) ** 5 / (1.0 - xi) ** 7) * (2.0 * math.pi) * alpha ** 3 / (n + 1)
    r = alpha * (1.0 + xi) / (1.0 - xi)
    return r[::-1].copy(), w[::-1].copy()


def atom_grid(z: int, acc: int, radial: str = "AHLRICHS"):
    """One atom's reference grid (points relative to the nucleus, weights incl. 4 pi r-quadrature)."""
    row = _row(z)
    n_rad = int(5.0 * (_RAD_ACC[acc - 1] + row - 8))
    if radial == "AHLRICHS":
        rp, rw = ahlrichs_radial(_AHLRICHS_ALPHA[z], n_rad)
    else:  # BECKE radial grid: alpha = Bragg-Slater radius (H) or half of it, in bohr
        from .geometry import ANGSTROM_TO_BOHR
        bs = _BRAGG_SLATER[z] * ANGSTROM_TO_BOHR
        rp, rw = becke_radial(bs if z == 1 else 0.5 * bs, n_rad)
    zone_row = 2 if row > 3 else row - 1
    redp1 = 2 if (row == 1 and acc > 1) else 1
    zones = [r * _CLEMENTI[z] for r in _RANGES[zone_row]]
    pts, wts = [], []
    sph_acc = 0
    for i in range(n_rad):
        if sph_acc < 4 and rp[i] > zones[sph_acc]:
            sph_acc += 1
        x, w = _lebedev(_LDVAL[acc - redp1][sph_acc])
        pts.append(x * rp[i])
        wts.append(w * rw[i] * 4.0 * math.pi)
    return np.concatenate(pts, axis=0), np.concatenate(wts)


def _load_helper():
    path = os.path.join(_HERE, "libsxc_inputs.so")
    if not os.path.exists(path):
        raise RuntimeError("serenity_b200/inputs/libsxc_inputs.so missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = ctypes.CDLL(path)
    lib.sxc_partition_weights.restype = None
    lib.sxc_partition_weights.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p,
                                          ctypes.c_void_p]
    return lib


_HELPER = None
_FLAVOURS = {"BECKE": 0, "SSF": 1, "VORONOI": 2}


def hilbert_index(ipts: np.ndarray, bits: int = 10) -> np.ndarray:
    """3-D Hilbert-curve index (Skilling's transpose algorithm, vectorised). ipts: [n,3] ints in [0, 2^bits)."""
    X = [ipts[:, 0].astype(np.uint32).copy(), ipts[:, 1].astype(np.uint32).copy(), ipts[:, 2].astype(np.uint32).copy()]
    M = np.uint32(1 << (bits - 1))
    Q = M
    while Q > 1:  # inverse undo excess work
        P = np.uint32(Q - 1)
        for i in range(3):
            hit = (X[i] & Q) != 0
            X[0] = np.where(hit, X[0] ^ P, X[0])
            t = np.where(hit, np.uint32(0), (X[0] ^ X[i]) & P)
            X[0] ^= t
            X[i] ^= t
        Q = np.uint32(Q >> 1)
    for i in range(1, 3):  # Gray encode
        X[i] ^= X[i - 1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ np.uint32(Q - 1), t)
        Q = np.uint32(Q >> 1)
    for i in range(3):
        X[i] ^= t
    idx = np.zeros(ipts.shape[0], dtype=np.uint64)
    for b in range(bits - 1, -1, -1):  # interleave, x most significant
        for i in range(3):
            idx = (idx << np.uint64(1)) | ((X[i] >> np.uint32(b)) & np.uint32(1)).astype(np.uint64)
    return idx


# lookup tables of the reference's Hilbert sort (src/grid/HilbertRTreeSorting.cpp:57-78): cube transformation to the next
# level of depth, and the number of a point inside the first cube from its relation to the centre
_HRT_TRANS = np.array([[0, 7, 6, 1, 2, 5, 4, 3], [0, 3, 4, 6, 7, 5, 2, 1], [0, 3, 4, 6, 7, 5, 2, 1], [2, 3, 0, 1, 6, 7, 4, 5],
                       [2, 3, 0, 1, 6, 7, 4, 5], [6, 5, 2, 1, 0, 3, 4, 7], [6, 5, 2, 1, 0, 3, 4, 7], [4, 3, 2, 5, 6, 1, 0, 7]])
_HRT_VAL = np.array([[[5, 6], [4, 7]], [[2, 1], [3, 0]]])


def hilbert_rtree_order(xyz: np.ndarray) -> np.ndarray:
    """The permutation HilbertRTreeSorting::sort applies (src/grid/HilbertRTreeSorting.cpp:29-214): depth from the number of
    points (:32-39), integer coordinates (pt - min) * length / extent truncated (:83-86), the index from the two lookup tables
    (:88-108), points moved in DESCENDING index order (:134-140 take the maximum first, the merge :171-181 keeps the larger
    one).  Ties keep their input order (what the reference does with one sorting node; with several nodes its order of
    equal indices depends on the thread count)."""
    xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
    n = xyz.shape[0]
    depth, length, nvert = 1, 2, 8
    while n * 8 > nvert:
        depth, length, nvert = depth + 1, length * 2, nvert * 8
    if depth > 20:
        raise ValueError("grid too large for a 63-bit Hilbert index")
    lo = xyz.min(axis=0)
    spread = float(length) / (xyz.max(axis=0) - lo)
    pt = ((xyz - lo) * spread).astype(np.int64)  # C++ int(): truncation
    l = length // 2
    g = pt > l
    v = _HRT_VAL[g[:, 0].astype(int), g[:, 1].astype(int), g[:, 2].astype(int)]
    idx = v.astype(np.int64)
    while l > 1:
        idx *= 8
        pt = pt - g * l
        l //= 2
        g = pt > l
        x = _HRT_VAL[g[:, 0].astype(int), g[:, 1].astype(int), g[:, 2].astype(int)]
        v = _HRT_TRANS[v, x]
        idx += v
    return np.argsort(-idx, kind="stable")


def becke_size_adjustments(zs):
    """a(j + nAtoms * i) of GridFactory.cpp:95-113 (Bragg-Slater radii), as a [nat, nat] C-contiguous array."""
    from .geometry import ANGSTROM_TO_BOHR
    nat = len(zs)
    bs = np.asarray([_BRAGG_SLATER[z] for z in zs]) * ANGSTROM_TO_BOHR
    aij = np.zeros((nat, nat))
    for i in range(nat):
        for j in range(nat):
            q = math.sqrt(bs[j] / bs[i])
            u = (q - 1.0) / (q + 1.0)
            a = u / (u * u - 1.0)
            aij[i, j] = min(0.5, max(-0.5, a))  # aOfAtomPair(j, i) stored at data()[j + nAtoms*i]
    return np.ascontiguousarray(aij)


def reference_atom_grids(symbols, coords_bohr, acc: int = 4, radial: str = "AHLRICHS"):
    """Every atom's reference grid shifted to its nucleus: (xyz [N, 3], atomic weights [N], parent atom [N])."""
    from .geometry import atomic_numbers
    zs = atomic_numbers(symbols)
    coords = np.ascontiguousarray(coords_bohr, dtype=np.float64)
    cache, pts, wts, par = {}, [], [], []
    for k, z in enumerate(zs):
        if z not in cache:
            cache[z] = atom_grid(z, acc, radial)
        p0, w0 = cache[z]
        pts.append(p0 + coords[k])
        wts.append(w0)
        par.append(np.full(w0.shape[0], k, dtype=np.int32))
    return np.ascontiguousarray(np.concatenate(pts, axis=0)), np.concatenate(wts), np.concatenate(par)


def host_partition_weights(flavour, zs, coords, xyz, w_atomic, parent, smoothing: int = 3):
    """CPU restatement of the weight step (gridweights.c), atom by atom; the checker of XCContext.partition_weights."""
    global _HELPER
    if _HELPER is None:
        _HELPER = _load_helper()
    nat = len(zs)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    adist = np.ascontiguousarray(np.linalg.norm(coords[:, None, :] - coords[None, :, :], axis=2))
    aij = becke_size_adjustments(zs) if flavour != "SSF" else None
    out = np.array(w_atomic, dtype=np.float64, copy=True)
    for k in range(nat):
        sel = np.nonzero(parent == k)[0]
        pts = np.ascontiguousarray(xyz[sel])
        w = np.ascontiguousarray(out[sel])
        _HELPER.sxc_partition_weights(_FLAVOURS[flavour], smoothing, nat, coords.ctypes.data, adist.ctypes.data,
                                      aij.ctypes.data if aij is not None else None, k, pts.shape[0],
                                      pts.ctypes.data, w.ctypes.data)
        out[sel] = w
    return out


def molecular_grid(symbols, coords_bohr, acc: int = 4, flavour: str = "SSF", radial: str = "AHLRICHS",
                   weight_threshold: float = 1e-14, sort: bool = True, device_ctx=None, smoothing: int = 3):
    """Returns (xyz [N,3] float64 C-contiguous == Matrix3Xd column-major interleaved, w [N]).

    device_ctx: an XCContext -> the O(N n_atoms^2) partition-weight step runs on the GPU (sxc_partition_weights);
    default: the host restatement (the synthetic inputs of tests and bench.py are made this way)."""
    from .geometry import atomic_numbers
    zs = atomic_numbers(symbols)
    coords = np.ascontiguousarray(coords_bohr, dtype=np.float64)
    xyz, w0, parent = reference_atom_grids(symbols, coords, acc, radial)
    if device_ctx is not None:
        w, _ = device_ctx.partition_weights(flavour, coords, xyz, parent, w0,
                                            becke_size_adjustments(zs) if flavour != "SSF" else None, smoothing)
    else:
        w = host_partition_weights(flavour, zs, coords, xyz, w0, parent, smoothing)
    keep = w > weight_threshold  # GridFactory.cpp:264
    xyz, w = xyz[keep], w[keep]
    if sort == "reference":  # the reference's own curve and order (HilbertRTreeSorting.cpp)
        order = hilbert_rtree_order(xyz)
        xyz, w = xyz[order], w[order]
    elif sort:
        lo = xyz.min(axis=0)
        span = np.maximum(xyz.max(axis=0) - lo, 1e-300)
        ip = np.minimum(((xyz - lo) / span * 1024.0).astype(np.int64), 1023)
        order = np.argsort(hilbert_index(ip, 10), kind="stable")
        xyz, w = xyz[order], w[order]
    return np.ascontiguousarray(xyz), np.ascontiguousarray(w)
