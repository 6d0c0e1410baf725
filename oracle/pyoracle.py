"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under serenity_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


_SRCS = ("oracle.c", "oracle_functionals.c", "oracle_functionals_u.cpp", "oracle_kernel2.cpp", "ozaki.c",
         "functionals_jet.inc", "oracle.h", "harmonics_table.h", "harmonics_gen.h", "Makefile")


def _cpu_has_avx512() -> bool:
    need = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"}
    try:
        with open("/proc/cpuinfo") as fh:
            for ln in fh:
                if ln.startswith("flags"):
                    return need <= set(ln.split(":", 1)[1].split())
    except OSError:
        pass
    return False


def variant() -> str:
    """'avx512' (x86-64-v4 build, 512-bit vectors) when the host CPU has it and ORACLE_VARIANT does not say otherwise, else 'avx2'."""
    want = os.environ.get("ORACLE_VARIANT", "")
    if want in ("avx2", "avx512"):
        return want
    return "avx512" if _cpu_has_avx512() else "avx2"


def _lib_path(var=None) -> str:
    return os.path.join(_HERE, "liboracle_avx512.so" if (var or variant()) == "avx512" else "liboracle.so")


def build(force: bool = False) -> str:
    """Both variants (they travel prebuilt to the GPU box); returns the portable one."""
    srcs = [os.path.join(_HERE, f) for f in _SRCS]
    for path in (_lib_path("avx2"), _lib_path("avx512")):
        if force or not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "-j4"])
            break
    return _lib_path("avx2")


_BLAS = None


def use_openblas(on: bool = True) -> dict:
    """Route the oracle's dense products (phi_s P_s, phi_s^T G) through OpenBLAS' cblas_dgemm from the scipy / numpy wheel,
    one thread per call (the OpenMP loop over blocks supplies the parallelism, as in the reference where Eigen's GEMM runs
    inside `omp parallel for`).  Returns {'gemm': ..., 'library': ..., 'core': ...}; falls back to the built-in loops."""
    global _BLAS
    L = lib()
    if not on:
        L.orc_set_dgemm(None)
        return {"gemm": "builtin loops (gcc auto-vectorised, %s)" % variant()}
    import glob
    import site
    cands = []
    for sp in site.getsitepackages() + [os.path.dirname(os.path.dirname(np.__file__))]:
        cands += sorted(glob.glob(os.path.join(sp, "scipy.libs", "libscipy_openblas-*.so")))
        cands += sorted(glob.glob(os.path.join(sp, "numpy.libs", "libscipy_openblas-*.so")))
    for path in cands:
        try:
            B = C.CDLL(path)
            fn = B.scipy_cblas_dgemm
        except (OSError, AttributeError):
            continue
        try:
            B.scipy_openblas_set_num_threads(1)
            B.scipy_openblas_get_corename.restype = C.c_char_p
            core = B.scipy_openblas_get_corename().decode()
        except AttributeError:
            core = "unknown"
        _BLAS = B
        L.orc_set_dgemm(C.cast(fn, C.c_void_p))
        return {"gemm": "OpenBLAS cblas_dgemm (LP64, 1 thread per call inside the OpenMP block loop)",
                "library": os.path.basename(path), "core": core}
    L.orc_set_dgemm(None)
    return {"gemm": "builtin loops (gcc auto-vectorised, %s): no loadable OpenBLAS with scipy_cblas_dgemm found" % variant()}


class _Basis(C.Structure):
    _fields_ = [("nshell", C.c_int), ("nbf", C.c_int), ("l", C.c_void_p), ("pure", C.c_void_p),
                ("nprim", C.c_void_p), ("prim_off", C.c_void_p), ("first_bf", C.c_void_p), ("centre", C.c_void_p),
                ("alpha", C.c_void_p), ("coeff", C.c_void_p), ("normfac", C.c_void_p)]


class _Grid(C.Structure):
    _fields_ = [("npts", C.c_long), ("xyz", C.c_void_p), ("w", C.c_void_p), ("blocksize", C.c_int)]


class _Func(C.Structure):
    _fields_ = [("ncomp", C.c_int), ("id", C.c_void_p), ("mix", C.c_void_p)]


class Timings(C.Structure):
    _fields_ = [("basis_on_grid", C.c_double), ("density_on_grid", C.c_double), ("functional", C.c_double),
                ("grid_to_matrix", C.c_double), ("total", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        path = _lib_path()
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.orc_set_dgemm.argtypes = [C.c_void_p]
        _LIB.orc_functional_on_grid.restype = C.c_double
        _LIB.orc_functional_on_grid_u.restype = C.c_double
        _LIB.orc_nblocks.restype = C.c_int
        _LIB.orc_max_threads.restype = C.c_int
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Basis:
    def __init__(self, tab):
        self.tab = tab
        self._keep = [np.ascontiguousarray(x) for x in (tab.l, tab.pure, tab.nprim, tab.prim_off, tab.first_bf,
                                                        tab.centre, tab.alpha, tab.coeff, tab.normfac)]
        k = self._keep
        self.c = _Basis(tab.nshell, tab.nbf, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]),
                        _p(k[7]), _p(k[8]))
        self.nbf = tab.nbf


class Grid:
    def __init__(self, xyz, w, blocksize=128):
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        self.c = _Grid(self.w.shape[0], _p(self.xyz), _p(self.w), blocksize)
        self.npts = self.w.shape[0]
        self.blocksize = blocksize

    @property
    def nblocks(self):
        return (self.npts + self.blocksize - 1) // self.blocksize


class Functional:
    def __init__(self, ids, mix):
        self.ids = np.asarray(ids, dtype=np.int32)
        self.mix = np.asarray(mix, dtype=np.float64)
        self.c = _Func(len(self.ids), _p(self.ids), _p(self.mix))

    @property
    def is_gga(self):
        return bool(lib().orc_functional_is_gga(C.byref(self.c)))


def max_threads():
    return lib().orc_max_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))


def basis_block(basis: Basis, grid: Grid, radial_thr: float, deriv: int, block: int):
    n = min(grid.blocksize, grid.npts - block * grid.blocksize)
    nb = basis.nbf
    arrs = [np.zeros((nb, n)) for _ in range(1 + (3 if deriv >= 1 else 0) + (6 if deriv >= 2 else 0))]
    ptrs = [_p(a) for a in arrs] + [None] * (10 - len(arrs))
    neg = np.zeros(nb, dtype=np.int32)
    cen = np.zeros(3)
    got = lib().orc_basis_block(C.byref(basis.c), C.byref(grid.c), C.c_double(radial_thr), deriv, block, *ptrs,
                                _p(neg), _p(cen))
    assert got == n
    # arrays are [nbf][n] (function-major == n x nbf column-major); return as [n, nbf] views
    return [a.T for a in arrs], neg, cen


def density_on_grid(basis: Basis, grid: Grid, radial_thr: float, P, deriv: int = 1):
    N = grid.npts
    P = np.asfortranarray(P, dtype=np.float64)
    rho = np.zeros(N)
    g = [np.zeros(N) for _ in range(3)] if deriv >= 1 else [None] * 3
    h = np.zeros((6, N)) if deriv >= 2 else None
    nonneg = np.zeros(grid.nblocks, dtype=np.int32)
    lib().orc_density_on_grid(C.byref(basis.c), C.byref(grid.c), C.c_double(radial_thr), _p(P), _p(rho), _p(g[0]),
                              _p(g[1]), _p(g[2]), _p(h), _p(nonneg))
    return rho, g, h, nonneg


def functional_on_grid(func: Functional, w, rho, gx=None, gy=None, gz=None):
    N = rho.shape[0]
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = [np.zeros(N) for _ in range(5)]
    gga = gx is not None
    e = lib().orc_functional_on_grid(C.byref(func.c), C.c_long(N), _p(w), _p(np.ascontiguousarray(rho)),
                                     _p(gx), _p(gy), _p(gz), _p(out[0]), _p(out[1]),
                                     _p(out[2]) if gga else None, _p(out[3]) if gga else None,
                                     _p(out[4]) if gga else None)
    return float(e), out


def basic_functional(fid: int, rho: float, sigma: float):
    F, a, s = C.c_double(), C.c_double(), C.c_double()
    rc = lib().orc_basic_functional(int(fid), C.c_double(rho), C.c_double(sigma), C.byref(F), C.byref(a), C.byref(s))
    if rc != 0:
        raise ValueError("unsupported functional id %d" % fid)
    return F.value, a.value, s.value


def scalar_to_matrix(basis: Basis, grid: Grid, radial_thr: float, block_ave_thr: float, v, gx=None, gy=None, gz=None):
    V = np.zeros((basis.nbf, basis.nbf), order="F")
    c = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (v, gx, gy, gz)]
    lib().orc_scalar_to_matrix(C.byref(basis.c), C.byref(grid.c), C.c_double(radial_thr), C.c_double(block_ave_thr),
                               _p(c[0]), _p(c[1]), _p(c[2]), _p(c[3]), _p(V))
    return V


def scalar_to_matrix_ab(basis_a: Basis, basis_b: Basis, grid: Grid, radial_thr: float, block_ave_thr: float, v, gx=None,
                        gy=None, gz=None):
    """Two-basis scatter (ScalarOperatorToMatrixAdder.cpp:216-220 / :286-300): [nbf_A, nbf_B]."""
    V = np.zeros((basis_a.nbf, basis_b.nbf), order="F")
    c = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (v, gx, gy, gz)]
    lib().orc_scalar_to_matrix_ab(C.byref(basis_a.c), C.byref(basis_b.c), C.byref(grid.c), C.c_double(radial_thr),
                                  C.c_double(block_ave_thr), _p(c[0]), _p(c[1]), _p(c[2]), _p(c[3]), _p(V))
    return V


def build_ab(basis_a: Basis, basis_b: Basis, densities, grid: Grid, func: Functional, radial_thr=1e-9, block_ave_thr=1e-11):
    """ABFuncPotential::getMatrix (ABFuncPotential.cpp:54-160), RESTRICTED: the densities of all (basis_C, P_C) pairs are
    summed on the grid, the functional is evaluated once, and its potential is scattered into the A x B matrix.
    Returns (V_AB, E_xc)."""
    N = grid.npts
    gga = func.is_gga
    tot = [np.zeros(N) for _ in range(4)]
    for bc, P in densities:
        rho, g, _, _ = density_on_grid(bc, grid, radial_thr, P, 1)
        tot[0] += rho
        for k in range(3):
            tot[1 + k] += g[k]
    e, out = functional_on_grid(func, grid.w, tot[0], *(tot[1:] if gga else (None,) * 3))
    V = scalar_to_matrix_ab(basis_a, basis_b, grid, radial_thr, block_ave_thr, out[1], *(out[2:5] if gga else (None,) * 3))
    return V, e


def build_ab_nadd(basis_a: Basis, basis_b: Basis, act, env, grid: Grid, func: Functional, radial_thr=1e-9, block_ave_thr=1e-11):
    """ABNAddFuncPotential::getMatrix (ABNAddFuncPotential.cpp:66-176), RESTRICTED: act = (Basis, P), env = [(Basis, P), ...];
    the potential v[rho_act + sum rho_env] - v[rho_act] (:150-170) scattered into the A x B matrix."""
    gga = func.is_gga
    ra, ga, _, _ = density_on_grid(act[0], grid, radial_thr, act[1], 1)
    tot = [ra.copy()] + [x.copy() for x in ga]
    for bc, P in env:
        rho, g, _, _ = density_on_grid(bc, grid, radial_thr, P, 1)
        tot[0] += rho
        for k in range(3):
            tot[1 + k] += g[k]
    _, sup = functional_on_grid(func, grid.w, tot[0], *(tot[1:] if gga else (None,) * 3))
    _, sub = functional_on_grid(func, grid.w, ra, *(ga if gga else (None,) * 3))
    v = [sup[k] - sub[k] for k in range(1, 5 if gga else 2)]
    return scalar_to_matrix_ab(basis_a, basis_b, grid, radial_thr, block_ave_thr, v[0], *(v[1:4] if gga else (None,) * 3))


def build_xc(basis: Basis, grid: Grid, func: Functional, P, radial_thr=1e-9, block_ave_thr=1e-11):
    P = np.asfortranarray(P, dtype=np.float64)
    V = np.zeros((basis.nbf, basis.nbf), order="F")
    E, ne = C.c_double(), C.c_double()
    t = Timings()
    rc = lib().orc_build_xc(C.byref(basis.c), C.byref(grid.c), C.byref(func.c), C.c_double(radial_thr),
                            C.c_double(block_ave_thr), _p(P), _p(V), C.byref(E), C.byref(ne), C.byref(t))
    if rc != 0:
        raise MemoryError("orc_build_xc failed")
    return V, E.value, ne.value, t


def build_nadd(basis_a: Basis, P_a, env, grid: Grid, func: Functional, radial_thr=1e-9, block_ave_thr=1e-11):
    """env: list of (Basis, P)."""
    P_a = np.asfortranarray(P_a, dtype=np.float64)
    Pe = [np.asfortranarray(p, dtype=np.float64) for _, p in env]
    nenv = len(env)
    barr = (C.POINTER(_Basis) * max(nenv, 1))(*[C.pointer(b.c) for b, _ in env])
    parr = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in Pe])
    V = np.zeros((basis_a.nbf, basis_a.nbf), order="F")
    E = C.c_double()
    parts = np.zeros(2 + nenv)
    rc = lib().orc_build_nadd(C.byref(basis_a.c), _p(P_a), nenv, barr, parr, C.byref(grid.c), C.byref(func.c),
                              C.c_double(radial_thr), C.c_double(block_ave_thr), _p(V), C.byref(E), _p(parts))
    if rc != 0:
        raise MemoryError("orc_build_nadd failed")
    return V, E.value, parts


# ---------------------------------------------------------------------------------------------- UNRESTRICTED
def basic_functional_u(fid: int, ra: float, rb: float, gaa: float, gab: float, gbb: float):
    F = C.c_double()
    d = np.zeros(5)
    rc = lib().orc_basic_functional_u(int(fid), C.c_double(ra), C.c_double(rb), C.c_double(gaa), C.c_double(gab),
                                      C.c_double(gbb), C.byref(F), _p(d))
    if rc != 0:
        raise ValueError("unsupported functional id %d" % fid)
    return F.value, d


def functional_on_grid_u(func: Functional, w, rho2, grad23=None):
    """rho2 [2, N]; grad23 [2, 3, N] or None -> E, epuv [N], dFdRho [2, N], dFdGrad [2, 3, N]."""
    rho2 = np.ascontiguousarray(rho2, dtype=np.float64)
    N = rho2.shape[1]
    g = None if grad23 is None else np.ascontiguousarray(grad23, dtype=np.float64)
    ep, vr, vg = np.zeros(N), np.zeros((2, N)), np.zeros((2, 3, N))
    e = lib().orc_functional_on_grid_u(C.byref(func.c), C.c_long(N), _p(np.ascontiguousarray(w, dtype=np.float64)),
                                       _p(rho2), _p(g), _p(ep), _p(vr), _p(vg) if g is not None else None)
    return float(e), ep, vr, vg


def build_xc_u(basis: Basis, grid: Grid, func: Functional, Pa, Pb, radial_thr=1e-9, block_ave_thr=1e-11):
    Pa, Pb = (np.asfortranarray(p, dtype=np.float64) for p in (Pa, Pb))
    Va, Vb = (np.zeros((basis.nbf, basis.nbf), order="F") for _ in range(2))
    E, ne = C.c_double(), C.c_double()
    rc = lib().orc_build_xc_u(C.byref(basis.c), C.byref(grid.c), C.byref(func.c), C.c_double(radial_thr),
                              C.c_double(block_ave_thr), _p(Pa), _p(Pb), _p(Va), _p(Vb), C.byref(E), C.byref(ne))
    if rc != 0:
        raise MemoryError("orc_build_xc_u failed")
    return (Va, Vb), E.value, ne.value


def build_nadd_u(basis_a: Basis, Pa_pair, env, grid: Grid, func: Functional, radial_thr=1e-9, block_ave_thr=1e-11):
    """env: list of (Basis, (P_alpha, P_beta))."""
    PAa, PAb = (np.asfortranarray(p, dtype=np.float64) for p in Pa_pair)
    Pea = [np.asfortranarray(p[0], dtype=np.float64) for _, p in env]
    Peb = [np.asfortranarray(p[1], dtype=np.float64) for _, p in env]
    nenv = len(env)
    barr = (C.POINTER(_Basis) * max(nenv, 1))(*[C.pointer(b.c) for b, _ in env])
    pa = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in Pea])
    pb = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in Peb])
    Va, Vb = (np.zeros((basis_a.nbf, basis_a.nbf), order="F") for _ in range(2))
    E = C.c_double()
    parts = np.zeros(2 + nenv)
    rc = lib().orc_build_nadd_u(C.byref(basis_a.c), _p(PAa), _p(PAb), nenv, barr, pa, pb, C.byref(grid.c),
                                C.byref(func.c), C.c_double(radial_thr), C.c_double(block_ave_thr), _p(Va), _p(Vb),
                                C.byref(E), _p(parts))
    if rc != 0:
        raise MemoryError("orc_build_nadd_u failed")
    return (Va, Vb), E.value, parts


def xc_gradient(basis: Basis, grid: Grid, func: Functional, P, atom_of_bf, natoms: int, radial_thr=1e-9):
    """P: matrix (RESTRICTED) or (P_alpha, P_beta) -> [natoms, 3]."""
    unres = isinstance(P, (tuple, list))
    Pa = np.asfortranarray(P[0] if unres else P, dtype=np.float64)
    Pb = np.asfortranarray(P[1], dtype=np.float64) if unres else None
    amap = np.ascontiguousarray(atom_of_bf, dtype=np.int32)
    grad = np.zeros((natoms, 3), order="F")
    rc = lib().orc_xc_gradient(C.byref(basis.c), C.byref(grid.c), C.byref(func.c), C.c_double(radial_thr),
                               2 if unres else 1, _p(Pa), _p(Pb), natoms, _p(amap), _p(grad))
    if rc != 0:
        raise MemoryError("orc_xc_gradient failed")
    return grad


def nadd_gradient(basis_a: Basis, P_a, env, grid: Grid, func: Functional, atom_of_bf, natoms: int, radial_thr=1e-9):
    """NAddFuncPotential::getGeomGradients (RESTRICTED): env = [(Basis, P), ...] -> [natoms, 3] of the active system."""
    Pa = np.asfortranarray(P_a, dtype=np.float64)
    amap = np.ascontiguousarray(atom_of_bf, dtype=np.int32)
    mats = [np.asfortranarray(P, dtype=np.float64) for _, P in env]
    bptr = (C.c_void_p * max(len(env), 1))(*[C.addressof(b.c) for b, _ in env])
    pptr = (C.c_void_p * max(len(env), 1))(*[m.ctypes.data for m in mats])
    grad = np.zeros((natoms, 3), order="F")
    rc = lib().orc_nadd_gradient(C.byref(basis_a.c), _p(Pa), len(env), bptr, pptr, C.byref(grid.c), C.byref(func.c),
                                 C.c_double(radial_thr), natoms, _p(amap), _p(grad))
    if rc != 0:
        raise MemoryError("orc_nadd_gradient failed")
    return grad


# ---------------------------------------------------------------------------------------------- row f-4: kernel
def basic_functional_d2(fid: int, ra: float, rb: float, gaa: float, gab: float, gbb: float):
    """F, d5, H[5, 5] w.r.t. (rho_a, rho_b, s_aa, s_ab, s_bb)."""
    F = C.c_double()
    d, h = np.zeros(5), np.zeros((5, 5))
    rc = lib().orc_basic_functional_d2(int(fid), C.c_double(ra), C.c_double(rb), C.c_double(gaa), C.c_double(gab),
                                       C.c_double(gbb), C.byref(F), _p(d), _p(h))
    if rc != 0:
        raise ValueError("unsupported functional id %d" % fid)
    return F.value, d, h


def kernel_set_screen(thr: float = 1e-8):
    """test hook: density screen of storeDerivatives (reference: hard-coded 1e-8)"""
    lib().orc_kernel_set_screen(C.c_double(thr))


def kernel_store_r(func: Functional, rho, grad3=None, sign=1.0, store_gga=True, store=None):
    """Kernel<RESTRICTED>::storeDerivatives: store [10, N] (or [1, N]) += sign * d2F, screened at rho < 1e-8."""
    rho = np.ascontiguousarray(rho, dtype=np.float64)
    N = rho.shape[0]
    if store is None:
        store = np.zeros((10 if store_gga else 1, N))
    g = [None] * 3 if grad3 is None else [np.ascontiguousarray(x, dtype=np.float64) for x in grad3]
    lib().orc_kernel_store_r(C.byref(func.c), C.c_long(N), _p(rho), _p(g[0]), _p(g[1]), _p(g[2]), C.c_double(sign),
                             int(store_gga), _p(store))
    return store


def kernel_store_u(func: Functional, rho2, grad23=None, sign=1.0, store_gga=True, store=None):
    """Kernel<UNRESTRICTED>::storeDerivatives: rho2 [2, N], grad23 [2, 3, N] -> store [33, N] (or [3, N])."""
    rho2 = np.ascontiguousarray(rho2, dtype=np.float64)
    N = rho2.shape[1]
    if store is None:
        store = np.zeros((33 if store_gga else 3, N))
    g = None if grad23 is None else np.ascontiguousarray(grad23, dtype=np.float64)
    lib().orc_kernel_store_u(C.byref(func.c), C.c_long(N), _p(rho2), _p(g), C.c_double(sign), int(store_gga), _p(store))
    return store


def kernel_contract(basis: Basis, grid: Grid, store, D, mode: int, gga: bool, resp=None, radial_thr=1e-9,
                    block_ave_thr=1e-11):
    """contractKernel + contractBlock for one trial vector: D [nb, nb] (modes 0, 1) or [2, nb, nb] (mode 2), each
    column-major; resp [4 * nspin, N] is added to (created if None)."""
    nspin = 2 if mode == 2 else 1
    Dm = np.stack([np.asfortranarray(d, dtype=np.float64).ravel(order="F") for d in (D if mode == 2 else [D])])
    if resp is None:
        resp = np.zeros((4 * nspin, grid.npts))
    store = np.ascontiguousarray(store, dtype=np.float64)
    lib().orc_kernel_contract(C.byref(basis.c), C.byref(grid.c), C.c_double(radial_thr), C.c_double(block_ave_thr),
                              int(mode), int(gga), _p(store), _p(Dm), _p(resp))
    return resp


def kernel_integrate(basis: Basis, grid: Grid, resp, gga: bool, nspin: int = 1, radial_thr=1e-9, block_ave_thr=1e-11):
    """numericalIntegration + F += F^T for one trial vector -> [nspin] matrices (a single one for nspin = 1)."""
    nb = basis.nbf
    F = np.zeros((nspin, nb * nb))
    resp = np.ascontiguousarray(resp, dtype=np.float64)
    lib().orc_kernel_integrate(C.byref(basis.c), C.byref(grid.c), C.c_double(radial_thr), C.c_double(block_ave_thr),
                               int(gga), int(nspin), _p(resp), _p(F))
    mats = [F[s].reshape(nb, nb, order="F") for s in range(nspin)]
    return mats[0] if nspin == 1 else mats


# ---------------------------------------------------------------------------------------------- INT8-slice reference
def ozaki_slice_rows(X, k: int):
    """rows of X [rows, cols] -> (slices int8 [k, rows, cols], exponents int32 [rows])"""
    X = np.ascontiguousarray(X, dtype=np.float64)
    rows, cols = X.shape
    S = np.zeros((k, rows, cols), dtype=np.int8)
    e = np.zeros(rows, dtype=np.int32)
    lib().orc_ozaki_slice_rows(_p(X), rows, cols, int(k), _p(S), _p(e))
    return S, e


def ozaki_matmul(A, B, k: int):
    """A [m, K] . B [n, K]^T with k INT8 slices per operand row: (C float64 [m, n], acc int32 [k, m, n])."""
    Sa, ea = ozaki_slice_rows(A, k)
    Sb, eb = ozaki_slice_rows(B, k)
    m, K = Sa.shape[1], Sa.shape[2]
    n = Sb.shape[1]
    acc = np.zeros((k, m, n), dtype=np.int32)
    lib().orc_ozaki_gemm_i32(_p(Sa), _p(Sb), int(k), m, n, K, _p(acc))
    C_ = np.zeros((m, n))
    lib().orc_ozaki_combine(_p(acc), int(k), m, n, _p(ea), _p(eb), _p(C_))
    return C_, acc
