"""ctypes binding of libonsas_cuda.so (include/onsas_cuda.h).

Fails loudly when the shared library is missing: there is no CPU fallback on the product path.
Build it with `python __graft_entry__.py build` or `make -C onsas.jl_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libonsas_cuda.so")

# status codes (include/onsas_cuda.h)
OK, ERR_INVALID_ARG, ERR_NEGATIVE_VOLUME, ERR_CUDA, ERR_NOT_READY, ERR_UNSUPPORTED, ERR_COMM, ERR_ALLOC, ERR_BREAKDOWN = range(9)
MAT_SVK, MAT_NEOHOOKEAN, MAT_ISOLINEAR = 0, 1, 2
STRAIN_ROTATED_ENGINEERING, STRAIN_GREEN = 0, 1
FAMILY_TET, FAMILY_TRUSS = 0, 1
PRECOND_NONE, PRECOND_JACOBI, PRECOND_TWO_LEVEL = 0, 1, 2
OPT_CG_MODE, OPT_ASM_MINBLOCKS, OPT_CG_CHECK_EVERY, OPT_CG_BLOCKS_PER_SM, OPT_CG_PROFILE, OPT_FORCE_MG = 1, 2, 3, 4, 5, 6
OPT_HOST_CHUNKS, OPT_GJ_BLOCKED, OPT_HOST_MID_WEIGHT, OPT_COARSE_RBM, OPT_COARSE_FUSED, OPT_HOST_STREAMS = 7, 8, 9, 10, 11, 12
OPT_REORDER, OPT_CG_SINGLE_REDUCTION, OPT_TRUSS_MINBLOCKS, OPT_HOST_GRAPH, OPT_COARSE_GLOBAL, OPT_CG_L2_PREFETCH = 13, 14, 15, 16, 17, 18


class StepInfo(C.Structure):
    _fields_ = [("norm_dU", C.c_double), ("norm_U", C.c_double), ("norm_r", C.c_double), ("norm_Fext", C.c_double),
                ("cg_iters", C.c_int64), ("cg_residual", C.c_double), ("cg_tol", C.c_double),
                ("ms_assemble", C.c_double), ("ms_solve", C.c_double)]


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_vp = C.c_void_p

# every symbol declared in include/onsas_cuda.h: name -> (restype, argtypes)
SIGNATURES = {
    "onsas_create": (C.c_int32, [C.c_int32, C.POINTER(_vp)]),
    "onsas_destroy": (C.c_int32, [_vp]),
    "onsas_last_error": (C.c_char_p, [_vp]),
    "onsas_version": (C.c_int32, []),
    "onsas_set_stream": (C.c_int32, [_vp, _vp]),
    "onsas_set_option": (C.c_int32, [_vp, C.c_int32, C.c_int64]),
    "onsas_set_nodes": (C.c_int32, [_vp, C.c_int64, C.c_int64, C.c_int32, _dp]),
    "onsas_set_materials": (C.c_int32, [_vp, C.c_int32, _i32p, _dp]),
    "onsas_set_tets": (C.c_int32, [_vp, C.c_int64, _vp, _vp]),
    "onsas_set_trusses": (C.c_int32, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int32]),
    "onsas_set_free_dofs": (C.c_int32, [_vp, C.c_int64, _i64p, C.c_int64]),
    "onsas_finalize_mesh": (C.c_int32, [_vp]),
    "onsas_set_U": (C.c_int32, [_vp, _dp]),
    "onsas_get_U": (C.c_int32, [_vp, _dp]),
    "onsas_set_Fext": (C.c_int32, [_vp, _dp]),
    "onsas_get_Fint": (C.c_int32, [_vp, _dp]),
    "onsas_get_Fext": (C.c_int32, [_vp, _dp]),
    "onsas_add_face_load": (C.c_int32, [_vp, C.c_int64, _i32p, C.c_int32, _dp, C.POINTER(C.c_int32)]),
    "onsas_add_nodal_load": (C.c_int32, [_vp, C.c_int64, _i32p, _dp, C.POINTER(C.c_int32)]),
    "onsas_apply_loads": (C.c_int32, [_vp, C.c_int32, _dp]),
    "onsas_clear_loads": (C.c_int32, [_vp]),
    "onsas_get_dU": (C.c_int32, [_vp, _dp]),
    "onsas_assemble": (C.c_int32, [_vp]),
    "onsas_assemble_host": (C.c_int32, [_vp, _dp, _dp]),
    "onsas_eval_elements": (C.c_int32, [_vp, C.c_int32, C.c_int64, C.c_int64, _dp, _dp, _dp, _dp]),
    "onsas_newton_step": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int64, C.POINTER(StepInfo)]),
    "onsas_step": (C.c_int32, [_vp, C.c_int32, C.c_double, C.c_double, C.c_int64, C.c_int32, C.POINTER(StepInfo)]),
    "onsas_pcg": (C.c_int32, [_vp, _dp, _dp, C.c_int32, C.c_double, C.c_double, C.c_int64, C.POINTER(C.c_int64),
                              C.POINTER(C.c_double)]),
    "onsas_spmv": (C.c_int32, [_vp, _dp, _dp]),
    "onsas_spmv_resident": (C.c_int32, [_vp]),
    "onsas_synchronize": (C.c_int32, [_vp]),
    "onsas_get_csr_size": (C.c_int32, [_vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "onsas_get_csr": (C.c_int32, [_vp, _i64p, _i32p, _dp]),
    "onsas_get_stress_strain": (C.c_int32, [_vp, C.c_int32, _dp, _dp]),
    "onsas_get_table_stats": (C.c_int32, [_vp, _i64p]),
    "onsas_get_cg_profile": (C.c_int32, [_vp, _i64p]),
    "onsas_create_multi": (C.c_int32, [_i32p, C.c_int32, C.POINTER(_vp)]),
    "onsas_device_count": (C.c_int32, [_vp]),
    "onsas_part_create": (C.c_int32, [C.c_int32, C.c_int64, _dp, C.c_int64, _vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp,
                                      C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "onsas_part_destroy": (C.c_int32, [_vp]),
    "onsas_part_sizes": (C.c_int32, [_vp, C.c_int32, _i64p]),
    "onsas_part_local_to_global": (C.c_int32, [_vp, C.c_int32, _i32p]),
    "onsas_part_local_elements": (C.c_int32, [_vp, C.c_int32, C.c_int32, _i64p]),
    "onsas_part_halo_plan": (C.c_int32, [_vp, C.c_int32, _vp, _vp, _vp, _vp, _vp]),
    "onsas_part_load": (C.c_int32, [_vp, C.c_int32, _vp, C.c_int32]),
    "onsas_comm_unique_id": (C.c_int32, [_vp]),
    "onsas_comm_init": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp]),
    "onsas_set_halo": (C.c_int32, [_vp, C.c_int32, _vp, _vp, _vp, _vp]),
    "onsas_p2p_export": (C.c_int32, [_vp, _vp, C.POINTER(C.c_int64)]),
    "onsas_p2p_import": (C.c_int32, [_vp, _vp, _i64p, _vp]),
}


def build(verbose: bool = False) -> str:
    """Compile libonsas_cuda.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return SO_PATH


_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"{SO_PATH} is missing: the CUDA extension has not been built "
                "(run `python __graft_entry__.py build`); onsas.jl_b200 has no CPU fallback")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
