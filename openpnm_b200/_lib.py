"""ctypes binding of libb200pnm.so (the C-ABI in include/b200pnm.h).

There is deliberately no fallback: if the shared object is missing or no CUDA
device is visible, everything here raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# B200_LIB=dbg selects the bounds-checking build (python -m openpnm_b200.build --bounds)
LIB_PATH = os.path.join(HERE, "libb200pnm_dbg.so" if os.environ.get("B200_LIB") == "dbg"
                        else "libb200pnm.so")

OK = 0
ERR_NAMES = {-1: "CUDA", -2: "ARG", -3: "INDEX", -4: "NONFINITE", -5: "TOO_LARGE",
             -6: "NCCL", -7: "SINGULAR"}
MAT_PURE, MAT_SYSTEM = 0, 1
SRC_LINEAR, SRC_POWER_LAW, SRC_STANDARD_KINETICS = 0, 1, 2
FMT_NONE, FMT_CSR, FMT_DIA = 0, 1, 2
FMT_NAMES = {0: "none", 1: "csr", 2: "dia"}
PRECOND_JACOBI, PRECOND_AMG = 0, 1
PRECOND_CODES = {"jacobi": 0, "amg": 1}


UPDATE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)   # b200_update_fn


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200pnm error {ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz_system", C.c_int64), ("cg_iters", C.c_int64),
                ("fmt", C.c_int), ("ndiags", C.c_int), ("t_setup_ms", C.c_double),
                ("t_solve_ms", C.c_double), ("relres", C.c_double),
                ("kernel_launches", C.c_int64)]


_p = C.c_void_p
_i64 = C.c_int64
_dbl = C.c_double
_pi64 = C.POINTER(C.c_int64)
_pdbl = C.POINTER(C.c_double)
_pint = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/b200pnm.h declares
SIGNATURES = {
    "b200_version": (C.c_int, []),
    "b200_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "b200_destroy": (C.c_int, [_p]),
    "b200_last_error": (C.c_char_p, [_p]),
    "b200_stream": (_p, [_p]),
    "b200_set_stream": (C.c_int, [_p, _p]),
    "b200_synchronize": (C.c_int, [_p]),
    "b200_build_laplacian": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _pi64]),
    "b200_build_laplacian2": (C.c_int, [_p, _p, _p, _i64, _i64, _pi64]),
    "b200_build_laplacian2_rows": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _pi64]),
    "b200_add_to_diagonal": (C.c_int, [_p, _p]),
    "b200_check_topology": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _pi64, _pi64]),
    "b200_cluster_labels": (C.c_int, [_p, _p, _i64, _i64, _p, _pi64]),
    "b200_slab_clusters": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _pi64]),
    "b200_network_health": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _pi64]),
    "b200_trim": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p, _p, _pi64, _pi64]),
    "b200_apply_bcs": (C.c_int, [_p, _p, _p, C.c_int, _pdbl, _pi64]),
    "b200_pure_diag_sum": (C.c_int, [_p, _pdbl, _pi64]),
    "b200_set_diag_mean": (C.c_int, [_p, _dbl]),
    "b200_load_coo": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _pi64]),
    "b200_load_csr": (C.c_int, [_p, _p, _p, _p, _i64, _i64]),
    "b200_set_b": (C.c_int, [_p, _p]),
    "b200_get_b": (C.c_int, [_p, _p]),
    "b200_csr_shape": (C.c_int, [_p, C.c_int, _pi64, _pi64]),
    "b200_export_csr": (C.c_int, [_p, C.c_int, _p, _p, _p]),
    "b200_spmv": (C.c_int, [_p, C.c_int, _p, _p]),
    "b200_set_format": (C.c_int, [_p, C.c_int]),
    "b200_pcg_prepare": (C.c_int, [_p, C.c_int]),
    "b200_pcg": (C.c_int, [_p, _p, _dbl, _dbl, _i64, _p, _pi64, _pdbl, _pint]),
    "b200_pcg_format": (C.c_int, [_p, _pint, _pint]),
    "b200_set_precond": (C.c_int, [_p, C.c_int]),
    "b200_amg_setup": (C.c_int, [_p, _pint]),
    "b200_amg_level_shape": (C.c_int, [_p, C.c_int, _pi64, _pi64]),
    "b200_amg_level_info": (C.c_int, [_p, C.c_int, _pi64]),
    "b200_amg_export_level": (C.c_int, [_p, C.c_int, _p, _p, _p, _p]),
    "b200_clear_sources": (C.c_int, [_p]),
    "b200_add_source": (C.c_int, [_p, C.c_int, _p, _p, _p, _p]),
    "b200_picard": (C.c_int, [_p, _p, _dbl, _i64, _dbl, _dbl, _dbl, _i64, _p, _pi64, _pint,
                              _pi64, _pint]),
    "b200_picard_cb": (C.c_int, [_p, _p, _dbl, _i64, _dbl, _dbl, _dbl, _i64, _p, _pi64, _pint,
                                 _pi64, _pint, _p, _p]),
    "b200_set_conductance_model": (C.c_int, [_p, C.c_int, _p, _i64, _i64, _p, C.c_int, C.c_int, _p,
                                             _p, C.c_int, _p, _p, C.c_int]),
    "b200_clear_conductance_model": (C.c_int, [_p]),
    "b200_update_conductance": (C.c_int, [_p, _p]),
    "b200_conductance_values": (C.c_int, [_p, _p]),
    "b200_transient_rk45": (C.c_int, [_p, _p, _p, _dbl, _dbl, _dbl, _dbl, _p, _i64, _i64, _p, _p,
                                      _pi64, _pi64, _pi64, _pint]),
    "b200_rate_pores": (C.c_int, [_p, _p, _p, _i64, _p]),
    "b200_rate_throats": (C.c_int, [_p, _p, _p, _p, _p, _i64, _p]),
    "b200_rate_throats2": (C.c_int, [_p, _p, _p, _p, _p, _i64, _p]),
    "b200_source_values": (C.c_int, [_p, C.c_int, _p, _p, _p, _p]),
    "b200_get_stats": (C.c_int, [_p, C.POINTER(Stats)]),
    "b200_time_kernel": (C.c_int, [_p, C.c_int, C.c_int, _pdbl]),
    "b200_operator_bytes": (C.c_int, [_p, _pi64, _pi64]),
    "b200_solve_coo_h": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _p, _p, _dbl, _dbl, _i64, _p,
                                   _pi64, _pdbl, _pint]),
    "b200_nccl_unique_id": (C.c_int, [_p]),
    "b200_comm_init": (C.c_int, [_p, C.c_int, C.c_int, _p]),
    "b200_set_slab": (C.c_int, [_p, _i64, _i64, _i64]),
    "b200_p2p_export": (C.c_int, [_p, _p]),
    "b200_p2p_attach": (C.c_int, [_p, _p, _p]),
    "b200_p2p_disable": (C.c_int, [_p]),
}

_LIB = None


def load():
    """Load libb200pnm.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m openpnm_b200.build` "
            "(or __graft_entry__.build()).  openpnm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)    # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def as_c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
