"""ctypes binding of libefv_b200.so (the C-ABI declared in include/efv_b200.h).

The library is built in-tree by `pyefvlib_b200/csrc/build.sh` (also run by `__graft_entry__.build()`).
There is NO fallback: if the shared object is missing, or no CUDA device is present when a compute entry point
is called, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libefv_b200.so")

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_intp = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol of include/efv_b200.h (tests/test_host.py checks this)
PROTOTYPES = {
    "efv_last_error": (C.c_char_p, []),
    "efv_set_device": (C.c_int, [C.c_int]),
    "efv_device_count": (C.c_int, [c_intp]),
    "efv_get_stream": (C.c_void_p, []),
    "efv_synchronize": (C.c_int, []),
    "efv_kernel_launch_count": (C.c_int64, []),
    "efv_field_copy_counters": (C.c_int, [c_i64p, c_i64p, c_i64p, c_i64p]),
    "efv_mesh_create": (C.c_void_p, [C.c_int, C.c_int64, c_f64p, C.c_int64, c_i64p, c_i32p, c_i32p, C.c_int]),
    "efv_mesh_create_uniform": (C.c_void_p, [C.c_int, C.c_int64, c_f64p, C.c_int64, C.c_int, c_i32p, c_i32p, C.c_int]),
    "efv_mesh_set_boundaries": (C.c_int, [C.c_void_p, C.c_int, c_i64p, C.c_int64, c_i64p, c_i32p]),
    "efv_mesh_destroy": (None, [C.c_void_p]),
    "efv_mesh_num_vertices": (C.c_int64, [C.c_void_p]),
    "efv_mesh_num_elements": (C.c_int64, [C.c_void_p]),
    "efv_mesh_num_inner_faces": (C.c_int64, [C.c_void_p]),
    "efv_mesh_num_facets": (C.c_int64, [C.c_void_p]),
    "efv_mesh_num_outer_faces": (C.c_int64, [C.c_void_p]),
    "efv_mesh_group_size": (C.c_int64, [C.c_void_p, C.c_int]),
    "efv_geometry_build": (C.c_int, [C.c_void_p]),
    "efv_geometry_release": (C.c_int, [C.c_void_p]),
    "efv_geometry_export_group": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_f64p, c_f64p, c_f64p]),
    "efv_vertex_volume_export": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_boundary_export": (C.c_int, [C.c_void_p, c_i64p, c_i32p, c_f64p, c_i32p, c_f64p]),
    "efv_system_create": (C.c_void_p, [C.c_void_p, C.c_int]),
    "efv_system_create_from_csr": (C.c_void_p, [C.c_int64, c_i64p, c_i32p]),
    "efv_system_set_values": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_system_destroy": (None, [C.c_void_p]),
    "efv_system_rows": (C.c_int64, [C.c_void_p]),
    "efv_system_nnz": (C.c_int64, [C.c_void_p]),
    "efv_system_graph_nnz": (C.c_int64, [C.c_void_p]),
    "efv_system_slotmap_size": (C.c_int64, [C.c_void_p]),
    "efv_system_max_row_length": (C.c_int, [C.c_void_p]),
    "efv_system_export_csr": (C.c_int, [C.c_void_p, c_i64p, c_i32p]),
    "efv_system_export_graph": (C.c_int, [C.c_void_p, c_i64p, c_i32p]),
    "efv_system_export_slotmap": (C.c_int, [C.c_void_p, c_i64p]),
    "efv_system_export_values": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_vec_upload": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "efv_vec_download": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "efv_heat_assemble": (C.c_int, [C.c_void_p, c_f64p, C.c_double, C.c_int]),
    "efv_elasticity_assemble": (C.c_int, [C.c_void_p, c_f64p, C.c_int]),
    "efv_biot_assemble": (C.c_int, [C.c_void_p, c_f64p, C.c_double, c_f64p]),
    "efv_bc_clear": (C.c_int, [C.c_void_p]),
    "efv_bc_dirichlet": (C.c_int, [C.c_void_p, c_i64p, c_f64p, C.c_int64]),
    "efv_bc_neumann": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]),
    "efv_boundary_outer_faces": (C.c_int64, [C.c_void_p, C.c_int, c_i64p]),
    "efv_bc_neumann_values": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_f64p, C.c_int64, C.c_double]),
    "efv_assemble_dirichlet_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "efv_apply_dirichlet": (C.c_int, [C.c_void_p, C.c_int]),
    "efv_build_rhs": (C.c_int, [C.c_void_p, C.c_int]),
    "efv_fs_mass_rhs": (C.c_int, [C.c_void_p, C.c_void_p]),
    "efv_fs_geo_rhs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "efv_fs_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_f64p]),
    "efv_comm_unique_id": (C.c_int, [C.c_char_p, C.c_char_p]),
    "efv_comm_init": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_char_p]),
    "efv_comm_destroy": (C.c_int, []),
    "efv_comm_peer_export": (C.c_int, [C.c_char_p]),
    "efv_comm_peer_import": (C.c_int, [C.c_char_p]),
    "efv_halo_stage_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "efv_comm_peer_release": (C.c_int, []),
    "efv_halo_stage_import": (C.c_int, [C.c_void_p, C.c_char_p, c_i64p, c_intp, c_i64p]),
    "efv_system_set_owned": (C.c_int, [C.c_void_p, C.c_int64]),
    "efv_system_set_halo": (C.c_int, [C.c_void_p, C.c_int, c_intp, c_i64p, c_i32p, c_i64p]),
    "efv_halo_exchange": (C.c_int, [C.c_void_p, C.c_int]),
    "efv_solve_pcg": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_int, c_intp, c_f64p]),
    "efv_system_set_preconditioner": (C.c_int, [C.c_void_p, C.c_int]),
    "efv_system_set_lattice": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "efv_amg_level_export": (C.c_int, [C.c_void_p, C.c_int, c_i64p, c_i32p, c_f64p, c_i32p]),
    "efv_amg_info": (C.c_int, [C.c_void_p, C.c_int, c_intp, c_i64p, c_i64p, c_f64p, c_f64p]),
    "efv_solve_bicgstab": (C.c_int, [C.c_void_p, C.c_double, C.c_int, c_intp, c_f64p]),
    "efv_spmv": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "efv_true_residual": (C.c_int, [C.c_void_p, C.c_int, c_f64p]),
    "efv_vec_sums": (C.c_int, [C.c_void_p, C.c_int, c_f64p, c_f64p]),
    "efv_symmetry_defect": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_max_abs_diff": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_advance": (C.c_int, [C.c_void_p]),
    "efv_download_xold_begin": (C.c_int, [C.c_void_p]),
    "efv_download_xold_end": (C.c_int, [C.c_void_p, c_f64p]),
    "efv_last_timings": (C.c_int, [C.c_void_p, c_f64p, c_f64p]),
}

_lib = None


class EfvError(Exception):
    pass


def load():
    """dlopen the CUDA library; raises if it was not built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise EfvError(f"{LIB_PATH} is missing: build it with pyefvlib_b200/csrc/build.sh "
                           "(or __graft_entry__.build()); pyefvlib_b200 has no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name, None)
            if fn is None:
                continue        # not yet implemented symbols are reported by tests/test_host.py
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise EfvError(load().efv_last_error().decode("utf-8", "replace"))


def check_ptr(p):
    if not p:
        raise EfvError(load().efv_last_error().decode("utf-8", "replace"))
    return p


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p)


def i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(c_i32p)


def has_gpu():
    try:
        n = C.c_int(0)
        load().efv_device_count(C.byref(n))
        return n.value > 0
    except Exception:
        return False
