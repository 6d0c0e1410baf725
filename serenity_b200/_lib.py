"""ctypes binding of libserenity_xc_b200.so (the C ABI of include/serenity_xc_b200.h).

The product path has no CPU fallback: a missing library or a missing CUDA device raises SerenityError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libserenity_xc_b200.so")

# every symbol include/serenity_xc_b200.h declares
SYMBOLS = [
    "sxc_create", "sxc_destroy", "sxc_last_error", "sxc_set_stream", "sxc_set_workspace_limit", "sxc_set_tile_cache", "sxc_set_timing", "sxc_set_p_ready_event", "sxc_set_grid",
    "sxc_set_grid_shard", "sxc_add_basis", "sxc_set_functional", "sxc_build_xc", "sxc_build_xc_device",
    "sxc_build_nadd", "sxc_build_nadd_device", "sxc_xc_gradient", "sxc_density_on_grid", "sxc_basis_on_grid",
    "sxc_functional_on_grid", "sxc_functional_on_grid_u", "sxc_scalar_to_matrix", "sxc_get_stats", "sxc_balance_ranges", "sxc_abi_version",
    "sxc_partition_weights", "sxc_last_partition_ms", "sxc_scalar_to_matrix_ab", "sxc_build_ab", "sxc_build_ab_nadd", "sxc_nadd_gradient",
    "sxc_kernel_create", "sxc_kernel_destroy", "sxc_kernel_add", "sxc_kernel_get", "sxc_kernel_num_arrays", "sxc_kernel_contract",
    "sxc_kernel_integrate", "sxc_kernel_sigma", "sxc_kernel_response_copy", "sxc_kernel_contract_device", "sxc_kernel_integrate_device",
    "sxc_shell_table_from_file", "sxc_shell_table_sizes", "sxc_shell_table_copy", "sxc_shell_table_free", "sxc_add_basis_from_table",
    "sxc_host_last_error", "sxc_debug_scatter_schedule",
    "sxc_comm_unique_id", "sxc_comm_init_rank", "sxc_comm_destroy", "sxc_comm_info", "sxc_release_grid", "sxc_release_basis",
    "sxc_release_functional", "sxc_host_alloc", "sxc_host_free", "sxc_build_nadd_multi", "sxc_build_nadd_multi_device",
    "sxc_debug_scatter_schedule2",
    "sxc_group_create", "sxc_group_destroy", "sxc_group_size", "sxc_group_ctx", "sxc_group_last_error", "sxc_group_set_grid",
    "sxc_group_add_basis", "sxc_group_set_functional", "sxc_group_release_grid", "sxc_group_release_basis", "sxc_group_build_xc",
    "sxc_group_build_nadd_multi", "sxc_group_xc_gradient", "sxc_basis_hessian_on_grid", "sxc_density_hessian_on_grid",
    "sxc_supersystem_density_on_grid", "sxc_atom_grid", "sxc_molecular_grid", "sxc_hilbert_rtree_order", "sxc_grid_points_size",
    "sxc_grid_points_xyz", "sxc_grid_points_weights", "sxc_grid_points_free", "sxc_grid_last_error", "sxc_set_output_slice",
]


class SerenityError(RuntimeError):
    """Mirror of src/misc/SerenityError.h:36 - what the adapter throws when the C ABI returns an error."""


KERNEL_SLOTS = ["k_screen", "k_basis", "k_density", "k_functional", "k_form_g", "k_scatter", "finish", "allreduce"]  # SXC_T_*


class Stats(C.Structure):
    _fields_ = [("npts", C.c_int64), ("nblocks", C.c_int64), ("sum_s", C.c_int64), ("sum_ns", C.c_int64),
                ("sum_ns2", C.c_int64), ("sum_ns2_padded", C.c_int64), ("sum_s2", C.c_int64), ("s_max", C.c_int64),
                ("nbf", C.c_int64), ("workspace_bytes", C.c_int64), ("nchunks", C.c_int32),
                ("kernel_launches", C.c_int32), ("ms_kernel", C.c_float * 8), ("n_kernel", C.c_int32 * 8),
                ("ms_total", C.c_float)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("ms_kernel", "n_kernel")}
        d["ms_kernel"] = {n: float(self.ms_kernel[i]) for i, n in enumerate(KERNEL_SLOTS)}
        d["n_kernel"] = {n: int(self.n_kernel[i]) for i, n in enumerate(KERNEL_SLOTS)}
        return d


_LIB = None


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise SerenityError("libserenity_xc_b200.so is not built (run __graft_entry__.build()); "
                            "there is no CPU fallback for the XC build")
    lib = C.CDLL(LIB_PATH)
    vp, i, d, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    ip = C.POINTER(C.c_int)
    lib.sxc_create.argtypes = [C.POINTER(vp), i]
    lib.sxc_destroy.argtypes = [vp]
    lib.sxc_destroy.restype = None
    lib.sxc_last_error.argtypes = [vp]
    lib.sxc_last_error.restype = C.c_char_p
    lib.sxc_set_stream.argtypes = [vp, vp]
    lib.sxc_set_workspace_limit.argtypes = [vp, i64]
    lib.sxc_set_timing.argtypes = [vp, i]
    lib.sxc_set_output_slice.argtypes = [vp, i, i]
    lib.sxc_set_tile_cache.argtypes = [vp, i]
    lib.sxc_set_p_ready_event.argtypes = [vp, vp]
    lib.sxc_set_grid.argtypes = [vp, i64, vp, vp, i, ip]
    lib.sxc_set_grid_shard.argtypes = [vp, i, i, i]
    lib.sxc_add_basis.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp, vp, d, ip]
    lib.sxc_set_functional.argtypes = [vp, i, vp, vp, ip]
    lib.sxc_build_xc.argtypes = [vp, i, i, i, i, vp, d, vp, C.POINTER(d), C.POINTER(d)]
    lib.sxc_build_xc_device.argtypes = [vp, i, i, i, i, vp, d, vp]
    lib.sxc_build_nadd.argtypes = [vp, i, i, i, i, vp, i, vp, vp, i, d, vp, vp]
    lib.sxc_build_nadd_device.argtypes = [vp, i, i, i, i, vp, i, vp, vp, i, d, vp]
    lib.sxc_xc_gradient.argtypes = [vp, i, i, i, i, vp, i, vp, vp]
    lib.sxc_density_on_grid.argtypes = [vp, i, i, vp, vp, vp, vp, vp]
    lib.sxc_basis_on_grid.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, ip]
    lib.sxc_basis_hessian_on_grid.argtypes = [vp, i, i, i, vp, vp, vp, vp, vp, vp, ip]
    lib.sxc_density_hessian_on_grid.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp, vp]
    lib.sxc_supersystem_density_on_grid.argtypes = [vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.sxc_atom_grid.argtypes = [i, i, i, C.POINTER(vp)]
    lib.sxc_molecular_grid.argtypes = [vp, i, vp, vp, i, i, i, i, d, i, C.POINTER(vp)]
    lib.sxc_hilbert_rtree_order.argtypes = [i64, vp, vp]
    lib.sxc_grid_points_size.argtypes = [vp]
    lib.sxc_grid_points_size.restype = i64
    lib.sxc_grid_points_xyz.argtypes = [vp]
    lib.sxc_grid_points_xyz.restype = vp
    lib.sxc_grid_points_weights.argtypes = [vp]
    lib.sxc_grid_points_weights.restype = vp
    lib.sxc_grid_points_free.argtypes = [vp]
    lib.sxc_grid_points_free.restype = None
    lib.sxc_grid_last_error.restype = C.c_char_p
    lib.sxc_functional_on_grid.argtypes = [vp, i, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(d)]
    lib.sxc_functional_on_grid_u.argtypes = [vp, i, i64, vp, vp, i, vp, vp, C.POINTER(d)]
    lib.sxc_scalar_to_matrix.argtypes = [vp, i, i, d, vp, vp, vp, vp, vp]
    lib.sxc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.sxc_balance_ranges.argtypes = [i, vp, i, vp]
    lib.sxc_partition_weights.argtypes = [vp, i, i, i, vp, vp, i64, vp, vp, vp]
    lib.sxc_last_partition_ms.argtypes = [vp]
    lib.sxc_nadd_gradient.argtypes = [vp, i, i, i, i, vp, i, vp, vp, i, vp, vp]
    lib.sxc_scalar_to_matrix_ab.argtypes = [vp, i, i, i, d, vp, vp, vp, vp, vp]
    lib.sxc_build_ab.argtypes = [vp, i, i, i, i, i, i, vp, vp, d, vp, vp]
    lib.sxc_build_ab_nadd.argtypes = [vp, i, i, i, i, i, i, vp, i, vp, vp, d, vp]
    lib.sxc_last_partition_ms.restype = d
    lib.sxc_kernel_create.argtypes = [vp, i, i, i, ip]
    lib.sxc_kernel_destroy.argtypes = [vp, i]
    lib.sxc_kernel_add.argtypes = [vp, i, i, d, i, vp, vp]
    lib.sxc_kernel_get.argtypes = [vp, i, vp]
    lib.sxc_kernel_num_arrays.argtypes = [vp, i]
    lib.sxc_kernel_contract.argtypes = [vp, i, i, i, vp, i, i, vp, i]
    lib.sxc_kernel_integrate.argtypes = [vp, i, i, vp]
    lib.sxc_kernel_response_copy.argtypes = [vp, i, i]
    lib.sxc_kernel_contract_device.argtypes = [vp, i, i, i, vp, i, i, vp, i]
    lib.sxc_kernel_integrate_device.argtypes = [vp, i, i, vp]
    lib.sxc_shell_table_from_file.argtypes = [C.c_char_p, C.c_char_p, i, vp, vp, i, C.POINTER(vp)]
    lib.sxc_shell_table_sizes.argtypes = [vp, ip, ip, ip]
    lib.sxc_shell_table_copy.argtypes = [vp] + [vp] * 9
    lib.sxc_shell_table_free.argtypes = [vp]
    lib.sxc_shell_table_free.restype = None
    lib.sxc_add_basis_from_table.argtypes = [vp, vp, d, ip]
    lib.sxc_host_last_error.restype = C.c_char_p
    lib.sxc_debug_scatter_schedule.argtypes = [i, vp, i]
    lib.sxc_kernel_sigma.argtypes = [vp, i, i, i, vp, i, i, vp, vp]
    lib.sxc_build_nadd_multi.argtypes = [vp, i, i, vp, i, i, vp, i, vp, vp, i, d, i, vp, vp]
    lib.sxc_build_nadd_multi_device.argtypes = [vp, i, i, vp, i, i, vp, i, vp, vp, i, d, i, vp]
    lib.sxc_debug_scatter_schedule2.argtypes = [i, i, vp, i]
    lib.sxc_comm_unique_id.argtypes = [vp]
    lib.sxc_comm_init_rank.argtypes = [vp, i, i, vp]
    lib.sxc_comm_destroy.argtypes = [vp]
    lib.sxc_comm_info.argtypes = [vp, ip, ip, C.POINTER(i64), ip]
    lib.sxc_release_grid.argtypes = [vp, i]
    lib.sxc_release_basis.argtypes = [vp, i]
    lib.sxc_release_functional.argtypes = [vp, i]
    lib.sxc_host_alloc.argtypes = [C.c_size_t]
    lib.sxc_host_alloc.restype = vp
    lib.sxc_host_free.argtypes = [vp]
    lib.sxc_host_free.restype = None
    _LIB = lib
    return lib
