"""Python host side over the C ABI: XCContext (handle owner) and mirrors of the reference's Potential classes.

Plumbing only - every number is produced by the CUDA library (serenity_b200/csrc).  The class and method names
follow the reference (src/potentials/FuncPotential.h:47-147, src/potentials/NAddFuncPotential.h:118-155) so that
the parity tests read like the reference's own tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SerenityError, Stats


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _spin_pack(pair):
    """(M_alpha, M_beta), each [nb, nb] -> one buffer [2, nb*nb] of two column-major matrices back to back."""
    a, b = (np.asarray(m, dtype=np.float64) for m in pair)
    return np.ascontiguousarray(np.stack([a.reshape(-1, order="F"), b.reshape(-1, order="F")]))


def _spin_unpack(buf):
    nb = int(round(np.sqrt(buf.shape[1])))
    return tuple(np.asfortranarray(buf[s].reshape(nb, nb, order="F")) for s in range(2))


def _take_grid_points(lib, handle):
    n = lib.sxc_grid_points_size(handle)
    xyz = np.ctypeslib.as_array(C.cast(lib.sxc_grid_points_xyz(handle), C.POINTER(C.c_double)), shape=(n, 3)).copy()
    w = np.ctypeslib.as_array(C.cast(lib.sxc_grid_points_weights(handle), C.POINTER(C.c_double)), shape=(n,)).copy()
    lib.sxc_grid_points_free(handle)
    return xyz, w


def atom_grid(nuclear_charge: int, accuracy: int, radial: str = "AHLRICHS"):
    """Row f-1 through the C ABI (sxc_atom_grid, host only): AtomGridFactory::produce - points relative to the nucleus [n, 3]
    and weights [n] of one atom's pruned reference grid."""
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.sxc_atom_grid(int(nuclear_charge), int(accuracy), {"AHLRICHS": 0, "BECKE": 1}[radial], C.byref(h))
    if rc != 0:
        raise _lib.SerenityError(lib.sxc_grid_last_error().decode())
    return _take_grid_points(lib, h)


def hilbert_rtree_order(xyz):
    """sxc_hilbert_rtree_order (host only): the permutation HilbertRTreeSorting::sort applies."""
    lib = _lib.load()
    xyz = _f64(xyz).reshape(-1, 3)
    order = np.zeros(xyz.shape[0], dtype=np.int64)
    rc = lib.sxc_hilbert_rtree_order(xyz.shape[0], _ptr(xyz), _ptr(order))
    if rc != 0:
        raise _lib.SerenityError(lib.sxc_grid_last_error().decode())
    return order


def shell_table_from_file(path: str, basis_label: str, symbols, coords_bohr, spherical: bool = True):
    """Row f-2 through the C ABI (sxc_shell_table_from_file): BasisFunctionProvider + Shell + extended indices of the
    reference for a geometry and a Turbomole-format basis file.  Returns (ShellTable, atom_of_bf); host only."""
    from .inputs.basis import ShellTable
    lib = _lib.load()
    coords = _f64(coords_bohr).reshape(-1, 3)
    names = (C.c_char_p * len(symbols))(*[s.encode() for s in symbols])
    h = C.c_void_p()
    rc = lib.sxc_shell_table_from_file(path.encode(), basis_label.encode(), len(symbols), names, _ptr(coords), 1 if spherical else 0,
                                       C.byref(h))
    if rc != 0:
        raise SerenityError(lib.sxc_host_last_error().decode())
    try:
        ns, npr, nbf = C.c_int(), C.c_int(), C.c_int()
        lib.sxc_shell_table_sizes(h, C.byref(ns), C.byref(npr), C.byref(nbf))
        ns, npr, nbf = ns.value, npr.value, nbf.value
        li, pu, npm, fb = (np.zeros(ns, dtype=np.int32) for _ in range(4))
        cen, al, co, nf, aob = np.zeros((ns, 3)), np.zeros(npr), np.zeros(npr), np.zeros(nbf), np.zeros(nbf, dtype=np.int32)
        lib.sxc_shell_table_copy(h, _ptr(li), _ptr(pu), _ptr(npm), _ptr(fb), _ptr(cen), _ptr(al), _ptr(co), _ptr(nf), _ptr(aob))
    finally:
        lib.sxc_shell_table_free(h)
    off = np.concatenate([[0], np.cumsum(npm)[:-1]]).astype(np.int32)
    return ShellTable(li, pu, npm, off, fb, cen, al, co, nf, nbf), aob


class XCContext:
    """One context per process and GPU (sxc_create)."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.sxc_create(C.byref(h), int(device))
        if rc != 0:
            raise SerenityError("sxc_create failed (status %d): no usable CUDA device %d - the XC build has no CPU "
                                "fallback" % (rc, device))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sxc_destroy(self._h)
            self._h = None
            for ptr in self.__dict__.pop("_pinned", []):
                self._lib.sxc_host_free(ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SerenityError("serenity_xc_b200 status %d: %s" % (rc, self._lib.sxc_last_error(self._h).decode()))

    # ---- inputs
    def set_stream(self, cuda_stream_ptr):
        self._check(self._lib.sxc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_workspace_limit(self, nbytes: int):
        self._check(self._lib.sxc_set_workspace_limit(self._h, int(nbytes)))

    def set_p_ready_event(self, cuda_event_ptr: int):
        self._check(self._lib.sxc_set_p_ready_event(self._h, C.c_void_p(cuda_event_ptr)))

    def set_tile_cache(self, on: bool):
        """phi / grad phi tiles stay resident between builds of the same (grid, basis) pair (sxc_set_tile_cache)."""
        self._check(self._lib.sxc_set_tile_cache(self._h, 1 if on else 0))

    def set_timing(self, on: bool):
        self._check(self._lib.sxc_set_timing(self._h, 1 if on else 0))

    def set_grid(self, xyz, w, blocksize: int = 128) -> int:
        xyz = _f64(xyz).reshape(-1, 3)
        w = _f64(w)
        g = C.c_int(-1)
        self._check(self._lib.sxc_set_grid(self._h, w.shape[0], _ptr(xyz), _ptr(w), blocksize, C.byref(g)))
        return g.value

    def set_grid_shard(self, grid: int, rank: int, world: int):
        self._check(self._lib.sxc_set_grid_shard(self._h, grid, rank, world))

    def release_grid(self, grid: int):
        self._check(self._lib.sxc_release_grid(self._h, grid))

    def release_basis(self, basis: int):
        self._check(self._lib.sxc_release_basis(self._h, basis))

    # ---- multi-GPU: the communicator lives in the library (NCCL); only the 128-byte id travels through the caller
    @staticmethod
    def _prefer_bundled_nccl():
        """A Python process usually also holds PyTorch, which loads the NCCL of the nvidia-nccl wheel; if the library bound the
        system libnccl.so.2 first, a later `import torch` would find the wrong version under that soname.  Name the wheel's file
        (SXC_NCCL_LIBRARY) so that both use the same copy."""
        import os
        import sys
        if os.environ.get("SXC_NCCL_LIBRARY"):
            return
        for base in sys.path:
            cand = os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["SXC_NCCL_LIBRARY"] = cand
                return

    @staticmethod
    def comm_unique_id() -> bytes:
        XCContext._prefer_bundled_nccl()
        buf = C.create_string_buffer(128)
        rc = _lib.load().sxc_comm_unique_id(buf)
        if rc != 0:
            raise SerenityError("sxc_comm_unique_id failed (status %d): NCCL (libnccl.so.2) not loadable" % rc)
        return buf.raw

    def comm_init_rank(self, rank: int, world: int, unique_id: bytes):
        """Every grid of this context becomes shard `rank` of `world`; builds return the sum over ranks (one all-reduce)."""
        XCContext._prefer_bundled_nccl()
        self._check(self._lib.sxc_comm_init_rank(self._h, int(rank), int(world), C.create_string_buffer(bytes(unique_id), 128)))

    def comm_info(self) -> dict:
        r, w, v, n = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._check(self._lib.sxc_comm_info(self._h, C.byref(r), C.byref(w), C.byref(n), C.byref(v)))
        return {"rank": r.value, "world": w.value, "collectives": n.value, "nccl_version": v.value}

    def pinned_array(self, shape, order="F"):
        """float64 array in page-locked memory (sxc_host_alloc); freed with the context."""
        n = int(np.prod(shape))
        ptr = self._lib.sxc_host_alloc(max(n, 1) * 8)
        if not ptr:
            raise SerenityError("sxc_host_alloc failed")
        self.__dict__.setdefault("_pinned", []).append(ptr)
        buf = (C.c_double * max(n, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order=order)

    def add_basis(self, tab, radial_threshold: float = 1e-9) -> int:
        arrs = [np.ascontiguousarray(a) for a in (tab.l, tab.pure, tab.nprim, tab.first_bf, tab.centre, tab.alpha,
                                                  tab.coeff, tab.normfac)]
        b = C.c_int(-1)
        self._check(self._lib.sxc_add_basis(self._h, tab.nshell, *[_ptr(a) for a in arrs], radial_threshold, C.byref(b)))
        return b.value

    def set_functional(self, ids, mix) -> int:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        mix = _f64(mix)
        f = C.c_int(-1)
        self._check(self._lib.sxc_set_functional(self._h, len(ids), _ptr(ids), _ptr(mix), C.byref(f)))
        return f.value

    # ---- hot path
    def build_xc(self, grid, basis, func, P, block_ave_threshold: float = 1e-11, nspin: int = 1):
        """RESTRICTED: P [nb, nb] -> V [nb, nb].  UNRESTRICTED (nspin = 2): P = (P_alpha, P_beta) -> (V_alpha, V_beta)."""
        if nspin == 2:
            P = _spin_pack(P)
            V = np.zeros_like(P)
        else:
            P = np.asfortranarray(P, dtype=np.float64)
            V = np.zeros(P.shape, order="F")
        E, ne = C.c_double(), C.c_double()
        self._check(self._lib.sxc_build_xc(self._h, grid, basis, func, nspin, _ptr(P), block_ave_threshold, _ptr(V),
                                           C.byref(E), C.byref(ne)))
        return (_spin_unpack(V) if nspin == 2 else V), E.value, ne.value

    def build_xc_into(self, grid, basis, func, P, V, block_ave_threshold: float = 1e-11, nspin: int = 1):
        """sxc_build_xc on caller-owned host buffers (column-major float64 arrays, pageable or pinned): no allocation on the
        way; V may be None on ranks of a communicator that do not need the matrix.  Returns (E, nelec)."""
        E, ne = C.c_double(), C.c_double()
        self._check(self._lib.sxc_build_xc(self._h, grid, basis, func, nspin, P.ctypes.data_as(C.c_void_p), block_ave_threshold,
                                           None if V is None else V.ctypes.data_as(C.c_void_p), C.byref(E), C.byref(ne)))
        return E.value, ne.value

    def build_xc_device(self, grid, basis, func, d_P_ptr: int, d_VEN_ptr: int, block_ave_threshold: float = 1e-11,
                        nspin: int = 1):
        self._check(self._lib.sxc_build_xc_device(self._h, grid, basis, func, nspin, C.c_void_p(d_P_ptr),
                                                  block_ave_threshold, C.c_void_p(d_VEN_ptr)))

    def build_nadd(self, grid, func, basis_act, P_act, basis_env, P_env, env_frozen: bool = False,
                   block_ave_threshold: float = 1e-11, nspin: int = 1):
        if nspin == 2:
            P_act = _spin_pack(P_act)
            P_env = [_spin_pack(p) for p in P_env]
            V = np.zeros_like(P_act)
        else:
            P_act = np.asfortranarray(P_act, dtype=np.float64)
            P_env = [np.asfortranarray(p, dtype=np.float64) for p in P_env]
            V = np.zeros(P_act.shape, order="F")
        nenv = len(P_env)
        be = np.ascontiguousarray(basis_env, dtype=np.int32)
        pp = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in P_env])
        E = np.zeros(2 + nenv)
        self._check(self._lib.sxc_build_nadd(self._h, grid, func, nspin, basis_act, _ptr(P_act), nenv, _ptr(be), pp,
                                             int(env_frozen), block_ave_threshold, _ptr(V), _ptr(E)))
        return (_spin_unpack(V) if nspin == 2 else V), E

    def build_nadd_multi(self, grid, funcs, basis_act, P_act, basis_env, P_env, env_frozen: int = 0, sum_matrices: bool = True,
                         block_ave_threshold: float = 1e-11, nspin: int = 1):
        """All non-additive functionals of one FDE iteration in one device pass (sxc_build_nadd_multi).  Returns
        (V, E): V = the summed matrix (sum_matrices) or a list of one matrix per functional; E [nfunc, 2 + nenv]."""
        if nspin == 2:
            P_act = _spin_pack(P_act)
            P_env = [_spin_pack(p) for p in P_env]
        else:
            P_act = np.asfortranarray(P_act, dtype=np.float64)
            P_env = [np.asfortranarray(p, dtype=np.float64) for p in P_env]
        nenv, nf = len(P_env), len(funcs)
        nmat = 1 if sum_matrices else nf
        V = np.zeros((nmat,) + ((2, P_act.shape[1]) if nspin == 2 else (P_act.size,)))
        fh = np.ascontiguousarray(funcs, dtype=np.int32)
        be = np.ascontiguousarray(basis_env, dtype=np.int32)
        pp = (C.c_void_p * max(nenv, 1))(*[p.ctypes.data for p in P_env])
        E = np.zeros((nf, 2 + nenv))
        self._check(self._lib.sxc_build_nadd_multi(self._h, grid, nf, _ptr(fh), nspin, basis_act, _ptr(P_act), nenv, _ptr(be), pp,
                                                   int(env_frozen), block_ave_threshold, 1 if sum_matrices else 0, _ptr(V),
                                                   _ptr(E)))
        nb = int(round(np.sqrt(P_act.shape[-1] if nspin == 2 else P_act.size)))
        mats = [(_spin_unpack(V[m]) if nspin == 2 else np.asfortranarray(V[m].reshape(nb, nb, order="F"))) for m in range(nmat)]
        return (mats[0] if sum_matrices else mats), E

    def build_nadd_device(self, grid, func, basis_act, d_P_act: int, basis_env, d_P_env, d_VE: int,
                          env_frozen: bool = False, block_ave_threshold: float = 1e-11, nspin: int = 1):
        nenv = len(d_P_env)
        be = np.ascontiguousarray(basis_env, dtype=np.int32)
        pp = (C.c_void_p * max(nenv, 1))(*d_P_env)
        self._check(self._lib.sxc_build_nadd_device(self._h, grid, func, nspin, basis_act, C.c_void_p(d_P_act), nenv,
                                                    _ptr(be), pp, int(env_frozen), block_ave_threshold,
                                                    C.c_void_p(d_VE)))

    def xc_gradient(self, grid, basis, func, P, atom_of_bf, natoms: int, nspin: int = 1):
        """FuncPotential::getGeomGradients: [natoms, 3] XC contribution to the nuclear gradient."""
        P = _spin_pack(P) if nspin == 2 else np.asfortranarray(P, dtype=np.float64)
        amap = np.ascontiguousarray(atom_of_bf, dtype=np.int32)
        grad = np.zeros((natoms, 3), order="F")
        self._check(self._lib.sxc_xc_gradient(self._h, grid, basis, func, nspin, _ptr(P), natoms, _ptr(amap), _ptr(grad)))
        return grad

    def nadd_gradient(self, grid, func, basis_act, P_act, basis_env, P_env, atom_of_bf, natoms: int, nspin: int = 1):
        """NAddFuncPotential::getGeomGradients: [natoms, 3] over the atoms of the active system."""
        if nspin == 2:
            Pa = _spin_pack(P_act)
            mats = [_spin_pack(p) for p in P_env]
        else:
            Pa = np.asfortranarray(P_act, dtype=np.float64)
            mats = [np.asfortranarray(p, dtype=np.float64) for p in P_env]
        amap = np.ascontiguousarray(atom_of_bf, dtype=np.int32)
        hb = np.ascontiguousarray(basis_env, dtype=np.int32)
        ptrs = (C.c_void_p * max(len(mats), 1))(*[m.ctypes.data for m in mats])
        grad = np.zeros((natoms, 3), order="F")
        self._check(self._lib.sxc_nadd_gradient(self._h, grid, func, nspin, basis_act, _ptr(Pa), len(mats), _ptr(hb), ptrs,
                                                natoms, _ptr(amap), _ptr(grad)))
        return grad

    # ---- stage level
    def density_on_grid(self, grid, basis, P, npts: int, gradient: bool = True):
        P = np.asfortranarray(P, dtype=np.float64)
        rho = np.zeros(npts)
        g = [np.zeros(npts) for _ in range(3)] if gradient else [None, None, None]
        self._check(self._lib.sxc_density_on_grid(self._h, grid, basis, _ptr(P), _ptr(rho), _ptr(g[0]), _ptr(g[1]),
                                                  _ptr(g[2])))
        return rho, g

    def basis_on_grid(self, grid, basis, block: int, nbf: int, blocksize: int = 128):
        arrs = [np.zeros((nbf, blocksize)) for _ in range(4)]
        neg = np.zeros(nbf, dtype=np.int32)
        n = C.c_int(0)
        self._check(self._lib.sxc_basis_on_grid(self._h, grid, basis, block, *[_ptr(a) for a in arrs], _ptr(neg),
                                                C.byref(n)))
        n = n.value
        # library layout: n x nbf column-major (index mu*n + p)
        return [a.reshape(-1)[: n * nbf].reshape(nbf, n).T for a in arrs], neg, n

    def supersystem_density_on_grid(self, grid, bases, Ps, npts: int, gradient: bool = True):
        """SupersystemDensityOnGridController: sum of the subsystem densities (and gradients) on the common grid."""
        Ps = [np.asfortranarray(P, dtype=np.float64) for P in Ps]
        rho = np.zeros(npts)
        g = [np.zeros(npts) for _ in range(3)] if gradient else [None, None, None]
        barr = (C.c_int * len(bases))(*bases)
        parr = (C.c_void_p * len(Ps))(*[P.ctypes.data for P in Ps])
        self._check(self._lib.sxc_supersystem_density_on_grid(self._h, grid, len(bases), barr, parr, _ptr(rho), _ptr(g[0]),
                                                              _ptr(g[1]), _ptr(g[2])))
        return rho, g

    def basis_hessian_on_grid(self, grid, basis, block: int, nbf: int, blocksize: int = 128):
        """xx, xy, xz, yy, yz, zz second derivatives of every basis function on one block, each [n, nbf]."""
        arrs = [np.zeros((nbf, blocksize)) for _ in range(6)]
        n = C.c_int(0)
        self._check(self._lib.sxc_basis_hessian_on_grid(self._h, grid, basis, block, *[_ptr(a) for a in arrs], C.byref(n)))
        n = n.value
        return [a.reshape(-1)[: n * nbf].reshape(nbf, n).T for a in arrs], n

    def density_hessian_on_grid(self, grid, basis, P, npts: int):
        """xx, xy, xz, yy, yz, zz second derivatives of the density, each [npts]."""
        P = np.asfortranarray(P, dtype=np.float64)
        h = [np.zeros(npts) for _ in range(6)]
        self._check(self._lib.sxc_density_hessian_on_grid(self._h, grid, basis, _ptr(P), *[_ptr(a) for a in h]))
        return h

    def functional_on_grid(self, func, w, rho, gx=None, gy=None, gz=None):
        N = rho.shape[0]
        out = [np.zeros(N) for _ in range(5)]
        e = C.c_double()
        gga = gx is not None
        self._check(self._lib.sxc_functional_on_grid(
            self._h, func, N, _ptr(_f64(w)), _ptr(_f64(rho)), _ptr(_f64(gx)) if gga else None,
            _ptr(_f64(gy)) if gga else None, _ptr(_f64(gz)) if gga else None, _ptr(out[0]), _ptr(out[1]),
            _ptr(out[2]) if gga else None, _ptr(out[3]) if gga else None, _ptr(out[4]) if gga else None, C.byref(e)))
        return e.value, out

    def functional_on_grid_u(self, func, w, rho2, grad23=None):
        """rho2 [2, N], grad23 [2, 3, N] or None -> E, epuv [N], dFdRho [2, N], dFdGrad [2, 3, N]."""
        rho2 = _f64(rho2)
        N = rho2.shape[1]
        dens8 = np.zeros((8, N))
        dens8[0], dens8[4] = rho2[0], rho2[1]
        if grad23 is not None:
            dens8[1:4], dens8[5:8] = grad23[0], grad23[1]
        ep, out8 = np.zeros(N), np.zeros((8, N))
        e = C.c_double()
        self._check(self._lib.sxc_functional_on_grid_u(self._h, func, N, _ptr(_f64(w)), _ptr(dens8),
                                                       0 if grad23 is None else 1, _ptr(ep), _ptr(out8), C.byref(e)))
        return e.value, ep, out8[[0, 4]], np.stack([out8[1:4], out8[5:8]])

    def scalar_to_matrix(self, grid, basis, nbf, v, gx=None, gy=None, gz=None, block_ave_threshold: float = 1e-11,
                         V=None):
        if V is None:
            V = np.zeros((nbf, nbf), order="F")
        gga = gx is not None
        self._check(self._lib.sxc_scalar_to_matrix(self._h, grid, basis, block_ave_threshold, _ptr(_f64(v)),
                                                   _ptr(_f64(gx)) if gga else None, _ptr(_f64(gy)) if gga else None,
                                                   _ptr(_f64(gz)) if gga else None, _ptr(V)))
        return V

    def scalar_to_matrix_ab(self, grid, basis_a, basis_b, nbf_a, nbf_b, v, gx=None, gy=None, gz=None,
                            block_ave_threshold: float = 1e-11, V=None):
        """Two-basis scatter (ScalarOperatorToMatrixAdder.cpp:216-220 / :286-300): V [nbf_a, nbf_b] is added to."""
        if V is None:
            V = np.zeros((nbf_a, nbf_b), order="F")
        gga = gx is not None
        self._check(self._lib.sxc_scalar_to_matrix_ab(self._h, grid, basis_a, basis_b, block_ave_threshold, _ptr(_f64(v)),
                                                      _ptr(_f64(gx)) if gga else None, _ptr(_f64(gy)) if gga else None,
                                                      _ptr(_f64(gz)) if gga else None, _ptr(V)))
        return V

    def build_ab(self, grid, func, basis_a, basis_b, nbf_a, nbf_b, basis_c, P_c, block_ave_threshold: float = 1e-11,
                 nspin: int = 1):
        """ABFuncPotential::getMatrix (ABFuncPotential.cpp:54-160).  basis_c / P_c: handles and density matrices whose
        densities are summed; nspin = 2: every P_c is a (P_alpha, P_beta) pair and (V_alpha, V_beta) is returned.
        Returns (V_ab [nbf_a, nbf_b], E_xc, integral of the density)."""
        if nspin == 2:
            mats = [_spin_pack(p) for p in P_c]
        else:
            mats = [np.asfortranarray(p, dtype=np.float64) for p in P_c]
        V = np.zeros((nspin, nbf_a * nbf_b))
        ptrs = (C.c_void_p * len(mats))(*[m.ctypes.data for m in mats])
        hb = np.ascontiguousarray(basis_c, dtype=np.int32)
        E = (C.c_double * 2)()
        self._check(self._lib.sxc_build_ab(self._h, grid, func, nspin, basis_a, basis_b, len(mats), _ptr(hb), ptrs,
                                           block_ave_threshold, _ptr(V), E))
        out = [np.asfortranarray(V[s].reshape(nbf_a, nbf_b, order="F")) for s in range(nspin)]
        return (tuple(out) if nspin == 2 else out[0]), E[0], E[1]

    def build_ab_nadd(self, grid, func, basis_a, basis_b, nbf_a, nbf_b, basis_act, P_act, basis_env, P_env,
                      block_ave_threshold: float = 1e-11, nspin: int = 1):
        """ABNAddFuncPotential::getMatrix (ABNAddFuncPotential.cpp:66-176): v[rho_act + sum rho_env] - v[rho_act] scattered into
        the A x B matrix; nspin = 2: every P is a (P_alpha, P_beta) pair and (V_alpha, V_beta) is returned."""
        pack = (lambda p: _spin_pack(p)) if nspin == 2 else (lambda p: np.asfortranarray(p, dtype=np.float64))
        Pa = pack(P_act)
        mats = [pack(p) for p in P_env]
        V = np.zeros((nspin, nbf_a * nbf_b))
        ptrs = (C.c_void_p * max(len(mats), 1))(*[m.ctypes.data for m in mats])
        hb = np.ascontiguousarray(basis_env, dtype=np.int32)
        self._check(self._lib.sxc_build_ab_nadd(self._h, grid, func, nspin, basis_a, basis_b, basis_act, _ptr(Pa), len(mats),
                                                _ptr(hb), ptrs, block_ave_threshold, _ptr(V)))
        out = [np.asfortranarray(V[s].reshape(nbf_a, nbf_b, order="F")) for s in range(nspin)]
        return tuple(out) if nspin == 2 else out[0]

    def partition_weights(self, flavour: str, coords, xyz, parent, w, aij=None, smoothing: int = 3):
        """GridFactory.cpp:139-266 on the device: molecular partition weights (BECKE / SSF) of the atoms' reference
        grids.  xyz [N, 3] points (already shifted to their nuclei), parent [N] atom index, w [N] atomic weights;
        returns the molecular weights (0 where the SSF screen removes a point) and the kernel time in ms."""
        coords = _f64(coords).reshape(-1, 3)
        xyz = _f64(xyz).reshape(-1, 3)
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        out = _f64(w).copy()
        a = None if aij is None else _f64(aij)
        self._check(self._lib.sxc_partition_weights(self._h, {"BECKE": 0, "SSF": 1, "VORONOI": 2}[flavour], int(smoothing),
                                                    coords.shape[0], _ptr(coords),
                                                    None if a is None else _ptr(a), xyz.shape[0], _ptr(xyz), _ptr(parent),
                                                    _ptr(out)))
        return out, float(self._lib.sxc_last_partition_ms(self._h))

    def molecular_grid(self, nuclear_charges, coords_bohr, accuracy: int = 4, flavour: str = "SSF", radial: str = "AHLRICHS",
                       smoothing: int = 3, weight_threshold: float = 1e-14, hilbert_sort: bool = True):
        """GridFactory::produce behind the C ABI (sxc_molecular_grid): atom grids, partition weights on the device, weight cut,
        Hilbert R-tree order.  Returns (xyz [N, 3], w [N]) ready for set_grid."""
        z = np.ascontiguousarray(nuclear_charges, dtype=np.int32)
        coords = _f64(coords_bohr).reshape(-1, 3)
        h = C.c_void_p()
        rc = self._lib.sxc_molecular_grid(self._h, len(z), _ptr(z), _ptr(coords), int(accuracy),
                                          {"BECKE": 0, "SSF": 1, "VORONOI": 2}[flavour], {"AHLRICHS": 0, "BECKE": 1}[radial],
                                          int(smoothing), float(weight_threshold), 1 if hilbert_sort else 0, C.byref(h))
        if rc != 0:
            raise _lib.SerenityError(self._lib.sxc_grid_last_error().decode())
        return _take_grid_points(self._lib, h)

    # ---- row f-4: LR-TDDFT kernel
    def kernel_create(self, grid: int, nspin: int = 1, gga: bool = True) -> int:
        k = C.c_int(-1)
        self._check(self._lib.sxc_kernel_create(self._h, grid, nspin, 1 if gga else 0, C.byref(k)))
        return k.value

    def kernel_destroy(self, kernel: int):
        self._check(self._lib.sxc_kernel_destroy(self._h, kernel))

    def kernel_add(self, kernel: int, func: int, basis_c, P_c, sign: float = 1.0, nspin: int = 1):
        """One storeDerivatives call (Kernel.cpp:476-683) on the summed density of the (basis, P) pairs; nspin = 2: every
        P is a (P_alpha, P_beta) pair."""
        mats = [_spin_pack(p) if nspin == 2 else np.asfortranarray(p, dtype=np.float64) for p in P_c]
        ptrs = (C.c_void_p * len(mats))(*[m.ctypes.data for m in mats])
        hb = np.ascontiguousarray(basis_c, dtype=np.int32)
        self._check(self._lib.sxc_kernel_add(self._h, kernel, func, float(sign), len(mats), _ptr(hb), ptrs))

    def kernel_get(self, kernel: int, npts: int):
        n = self._lib.sxc_kernel_num_arrays(self._h, kernel)
        self._check(min(n, 0))
        out = np.zeros((n, npts))
        self._check(self._lib.sxc_kernel_get(self._h, kernel, _ptr(out)))
        return out

    @staticmethod
    def _pack_vectors(D, mode):
        """D: list of matrices (modes 0, 1) or of (D_alpha, D_beta) pairs (mode 2) -> [nvec * nspin, nb * nb]."""
        rows = []
        for d in D:
            for m in (d if mode == 2 else [d]):
                rows.append(np.asarray(m, dtype=np.float64).reshape(-1, order="F"))
        return np.ascontiguousarray(np.stack(rows))

    def kernel_contract(self, grid: int, basis_j: int, kernels, D, mode: int = 0, accumulate: bool = False):
        """contractKernel + contractBlock (KernelSigmavector.cpp:254-311, :360-497) for the trial densities D."""
        buf = self._pack_vectors(D, mode)
        hk = np.ascontiguousarray(kernels, dtype=np.int32)
        self._check(self._lib.sxc_kernel_contract(self._h, grid, basis_j, len(hk), _ptr(hk), mode, len(D), _ptr(buf),
                                                  1 if accumulate else 0))

    def kernel_sigma_device(self, grid: int, basis: int, kernels, d_D_ptr: int, d_F_ptr: int, nvec: int, mode: int = 0):
        """contract + integrate with device-resident D (nvec x nspin x nb^2) and F; asynchronous on the context's stream."""
        hk = np.ascontiguousarray(kernels, dtype=np.int32)
        self._check(self._lib.sxc_kernel_contract_device(self._h, grid, basis, len(hk), _ptr(hk), mode, nvec,
                                                         C.c_void_p(d_D_ptr), 0))
        self._check(self._lib.sxc_kernel_integrate_device(self._h, grid, basis, C.c_void_p(d_F_ptr)))

    def kernel_response_copy(self, grid: int, save: bool):
        """save: keep the current contracted response (the supersystem contraction); not save: restore it."""
        self._check(self._lib.sxc_kernel_response_copy(self._h, grid, 1 if save else 0))

    def kernel_integrate(self, grid: int, basis_i: int, nbf: int, nvec: int, mode: int = 0):
        """numericalIntegration + F += F^T (KernelSigmavector.cpp:313-358, :236-249) -> list of matrices / spin pairs."""
        nspin = 2 if mode == 2 else 1
        F = np.zeros((nvec * nspin, nbf * nbf))
        self._check(self._lib.sxc_kernel_integrate(self._h, grid, basis_i, _ptr(F)))
        mats = [np.asfortranarray(F[m].reshape(nbf, nbf, order="F")) for m in range(nvec * nspin)]
        return mats if nspin == 1 else [tuple(mats[2 * v:2 * v + 2]) for v in range(nvec)]

    def kernel_sigma(self, grid: int, basis: int, nbf: int, kernels, D, mode: int = 0):
        self.kernel_contract(grid, basis, kernels, D, mode, False)
        return self.kernel_integrate(grid, basis, nbf, len(D), mode)

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.sxc_get_stats(self._h, C.byref(s)))
        return s.as_dict()


# ------------------------------------------------------------------------------------------------------------------
# Mirrors of the reference's controllers / potentials (host logic only: lazy evaluation + notification)
# ------------------------------------------------------------------------------------------------------------------
class DensityMatrixController:
    """src/data/matrices/DensityMatrixController: owns P and notifies dependants when it changes."""

    def __init__(self, P):
        self._P = np.asfortranarray(P, dtype=np.float64)
        self._sensitive = []

    def addSensitiveObject(self, obj):
        self._sensitive.append(obj)

    def getDensityMatrix(self):
        return self._P

    def setDensityMatrix(self, P):
        self._P = np.asfortranarray(P, dtype=np.float64)
        for o in self._sensitive:
            o.notify()


class FuncPotential:
    """Drop-in body of FuncPotential<RESTRICTED> (src/potentials/FuncPotential.cpp:40-111).

    getMatrix() is lazy: recomputed only after notify() (density or grid changed, FuncPotential.h:107-109);
    getEnergy(P) returns the cached E_xc irrespective of P (FuncPotential.cpp:67-71).
    """

    def __init__(self, ctx: XCContext, grid: int, basis: int, dmat: DensityMatrixController, functional: int,
                 block_ave_threshold: float = 1e-11):
        self._ctx, self._grid, self._basis, self._dmat, self._func = ctx, grid, basis, dmat, functional
        self._thr = block_ave_threshold
        self._potential = None
        self._energy = 0.0
        self.n_electrons_on_grid = None
        dmat.addSensitiveObject(self)

    def notify(self):
        self._potential = None

    def getMatrix(self):
        if self._potential is None:
            V, E, ne = self._ctx.build_xc(self._grid, self._basis, self._func, self._dmat.getDensityMatrix(), self._thr)
            self._potential, self._energy, self.n_electrons_on_grid = V, E, ne
        return self._potential

    def getEnergy(self, P=None):
        if self._potential is None:
            self.getMatrix()
        return self._energy


class NAddFuncPotential:
    """Drop-in body of NAddFuncPotential<RESTRICTED> (src/potentials/NAddFuncPotential.cpp:192-326)."""

    def __init__(self, ctx: XCContext, grid: int, basis_act: int, act_dmat: DensityMatrixController, basis_env,
                 env_dmats, functional: int, block_ave_threshold: float = 1e-11):
        self._ctx, self._grid, self._func, self._thr = ctx, grid, functional, block_ave_threshold
        self._bA, self._dA, self._bE, self._dE = basis_act, act_dmat, list(basis_env), list(env_dmats)
        self._potential = None
        self._energy = 0.0
        self._env_frozen = False
        act_dmat.addSensitiveObject(self)
        for d in self._dE:
            d.addSensitiveObject(_EnvWatcher(self))

    def notify(self):
        self._potential = None

    def _env_changed(self):
        self._potential = None
        self._env_frozen = False

    def getMatrix(self):
        if self._potential is None:
            V, E = self._ctx.build_nadd(self._grid, self._func, self._bA, self._dA.getDensityMatrix(), self._bE,
                                        [d.getDensityMatrix() for d in self._dE], self._env_frozen, self._thr)
            self._potential = V
            self.energy_parts = E
            self._energy = E[0] - E[1] - E[2:].sum()  # NAddFuncPotential.cpp:249, :282-286
            self._env_frozen = True  # environment densities stay cached until one of them changes
        return self._potential

    def getEnergy(self, P=None):
        if self._potential is None:
            self.getMatrix()
        return self._energy


class _EnvWatcher:
    def __init__(self, owner):
        self._o = owner

    def notify(self):
        self._o._env_changed()
