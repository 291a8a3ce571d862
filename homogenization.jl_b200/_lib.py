"""ctypes binding of libhmg_b200.so (include/hmg.h).  No fallback: a missing library is an error."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HMG_LIB: a variant build of the same library (kernel experiments, tools/); the product is the in-tree default
LIB_PATH = os.environ.get("HMG_LIB") or os.path.join(_HERE, "libhmg_b200.so")

HMG_X, HMG_B, HMG_R, HMG_P, HMG_AP, HMG_V, HMG_W = range(7)
VEC_IDS = {"x": HMG_X, "b": HMG_B, "r": HMG_R, "p": HMG_P, "Ap": HMG_AP, "v": HMG_V, "w": HMG_W}

_i32, _i64, _f64 = C.c_int, C.c_int64, C.c_double
_p = C.c_void_p
_pd = C.POINTER(C.c_double)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/hmg.h
PROTOTYPES = {
    "hmg_last_error": (C.c_char_p, []),
    "hmg_version": (_i32, []),
    "hmg_create": (_i32, [_i32, _i32, _i64, _i64, _p, _p, _p, _f64, _i32, C.POINTER(_p)]),
    "hmg_destroy": (_i32, [_p]),
    "hmg_create_partitioned": (_i32, [_i32, _i32, _i64, _i64, _p, _p, _p, _f64, _i32, _i32, _i32, _p, _p, C.POINTER(_p)]),
    "hmg_nccl_unique_id": (_i32, [_p]),
    "hmg_nf": (_i64, [_p, _i32]),
    "hmg_ne_local": (_i64, [_p]),
    "hmg_ld": (_i64, [_p, _i32]),
    "hmg_group_width": (_i32, [_p]),
    "hmg_comm_mode": (_i32, [_p]),
    "hmg_local_elements": (_i32, [_p, _p]),
    "hmg_set_lambda": (_i32, [_p, _f64]),
    "hmg_set_sigma": (_i32, [_p, _p]),
    "hmg_upload": (_i32, [_p, _i32, _i32, _p, _i64]),
    "hmg_download": (_i32, [_p, _i32, _i32, _p, _i64]),
    "hmg_download_rows": (_i32, [_p, _i32, _i32, _i64, _p, _i64]),
    "hmg_copy_columns_from": (_i32, [_p, _i32, _i32, _p, _i32]),
    "hmg_fill": (_i32, [_p, _i32, _i32, _f64]),
    "hmg_copy": (_i32, [_p, _i32, _i32, _i32]),
    "hmg_axpy": (_i32, [_p, _i32, _f64, _i32, _i32]),
    "hmg_dot": (_i32, [_p, _i32, _i32, _i32, _pd]),
    "hmg_mul": (_i32, [_p, _i32, _f64, _i32, _i32]),
    "hmg_apply_global": (_i32, [_p, _i32, _i32, _i32]),
    "hmg_apply_constraint": (_i32, [_p, _i32, _i32]),
    "hmg_broadcast_interfaces": (_i32, [_p, _i32, _i32]),
    "hmg_zero_out_all_but_one": (_i32, [_p, _i32, _i32]),
    "hmg_local_residual": (_i32, [_p, _i32]),
    "hmg_restrict": (_i32, [_p, _i32]),
    "hmg_interpolate_add": (_i32, [_p, _i32]),
    "hmg_smoothing_steps": (_i32, [_p, _i32, _i32]),
    "hmg_set_coarse_matrix": (_i32, [_p, _i64, _p, _p, _p, _p]),
    "hmg_assemble_coarse": (_i32, [_p]),
    "hmg_copy_to_base": (_i32, [_p, _i32, _p]),
    "hmg_distribute": (_i32, [_p, _i32, _p]),
    "hmg_vcycle": (_i32, [_p, _i32, _i32, _pd]),
    "hmg_vcycles": (_i32, [_p, _i32, _i32, _i32, _p]),
    "hmg_rhs_axi_grad": (_i32, [_p, _p, _i32]),
    "hmg_integrate_first_term": (_i32, [_p, _i32, _p, _i64, _pd]),
    "hmg_integrate_terms": (_i32, [_p, _i32, _i32, _i64, _pd]),
    "hmg_integrate_area": (_i32, [_p, _i64, _pd]),
    "hmg_next_rhs": (_i32, [_p, _i32, _i32]),
    "hmg_generate_field": (_i32, [_i32, _p, C.c_uint64, _f64, _f64, _i32, _p, _p, _i32]),
    "hmg_refined_mesh": (_i32, [_p, _i32, _p, _p, C.POINTER(_i64)]),
    "hmg_synchronize": (_i32, [_p]),
    "hmg_time_op": (_i32, [_p, _i32, _i32, _i32, _i32, C.POINTER(C.c_float)]),
    "hmg_launch_count": (_i64, [_p]),
    "hmg_device_ptr": (_p, [_p, _i32, _i32]),
    "hmg_hier_to_lattice": (_i32, [_p, _i32, _p]),
}

# include/hmg_introspect.h (host-only verification entry points)
HOST_PROTOTYPES = {
    "hmg_host_last_error": (C.c_char_p, []),
    "hmg_host_reference": (_i32, [_i32, _i32, _i32, _p, _p, _p, _pd]),
    "hmg_host_refined_mesh": (_i32, [_i32, _i32, _i32, _p, _p, C.POINTER(_i64)]),
    "hmg_host_local_matrix": (_i32, [_i32, _i32, _i32, _p, _p]),
    "hmg_host_transfer_matrix": (_i32, [_i32, _i32, _i32, _p]),
    "hmg_host_interface_rows": (_i32, [_i32, _i32, _i32, _i32, _i32, _p, C.POINTER(_i64)]),
    "hmg_host_topology": (_i32, [_i32, _i64, _i64, _p, _i32, C.POINTER(_i64), C.POINTER(_i64), _p, _p, _p]),
    "hmg_host_boundary": (_i32, [_i32, _i64, _i64, _p, _p, _p]),
    "hmg_host_class_of": (_i32, [_i32, _i32, _i32]),
    "hmg_host_partition_elements": (_i32, [_i32, _i64, _i64, _p, _p, _i32, _i32, C.POINTER(_i64), _p, _p, _p, _p, _p]),
    "hmg_host_partition_peers": (_i32, [_i32, _i64, _i64, _p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "hmg_host_partition_cells": (_i32, [_i32, _i64, _i64, _p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "hmg_host_element_coefficients": (_i32, [_i32, _i64, _i64, _p, _p, _p, _p, _i32]),
    "hmg_host_apply_sweep": (_i32, [_i32, _i32, _i32, _i32, _p, _p, _p, _p]),
}

_lib = None


class HmgError(RuntimeError):
    pass


def _point_at_torch_nccl():
    """Partitioned contexts bind NCCL at run time; make them pick the copy PyTorch ships (the one a
    torchrun process has loaded anyway) unless the caller chose another through HMG_NCCL_LIB."""
    if os.environ.get("HMG_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["HMG_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def load():
    """Load libhmg_b200.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
            "homogenization.jl_b200 has no CPU fallback.")
    _point_at_torch_nccl()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in list(PROTOTYPES.items()) + list(HOST_PROTOTYPES.items()):
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise HmgError(load().hmg_last_error().decode("utf-8", "replace"))


def check_host(status):
    if status != 0:
        raise HmgError(load().hmg_host_last_error().decode("utf-8", "replace"))
