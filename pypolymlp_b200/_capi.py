"""ctypes binding of the C ABI in include/polymlp_b200.h (libpolymlp_b200.so).

The library is the product: there is no Python/NumPy fallback for any compute entry point.
If the shared library is missing, importing a compute class raises immediately.
"""

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpolymlp_b200.so")
DATA_DIR = os.path.join(_HERE, "data")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)

PM_FLAG_SIMPLE_KERNELS = 1
PM_FLAG_SCATTER = 2


class FeatureParamsC(C.Structure):
    _fields_ = [
        ("n_type", C.c_int), ("n_fn", C.c_int), ("pair_params", _dp), ("cond_offsets", _ip),
        ("cond_values", _ip), ("cutoff", C.c_double), ("model_type", C.c_int), ("max_p", C.c_int),
        ("max_l", C.c_int), ("n_lcomb", C.c_int), ("lcomb_order", _ip), ("l_comb", _ip),
        ("n_terms", _ip), ("lm_seq", _ip), ("lm_coeffs", _dp), ("feature_type", C.c_int),
    ]


class StructuresC(C.Structure):
    _fields_ = [
        ("n_st", C.c_int), ("axis", _dp), ("positions_c", _dp), ("types", _ip), ("n_atoms", _ip),
        ("force", _ip),
    ]


_lib = None

EXPORTS = [
    "pm_last_error", "pm_version", "pm_gtinv_read", "pm_model_create", "pm_model_destroy",
    "pm_model_n_features", "pm_model_info", "pm_model_type_info", "pm_model_polynomial", "pm_model_feature_attrs",
    "pm_model_count_flops", "pm_device_count", "pm_context_create", "pm_context_destroy",
    "pm_batch_rows", "pm_neighbor_full", "pm_features_x", "pm_fit_reset", "pm_fit_accumulate",
    "pm_fit_stage", "pm_fit_accumulate_staged", "pm_fit_accumulator", "pm_fit_fpad",
    "pm_fit_finalize", "pm_fit_finalize_view", "pm_fit_solve_ridge", "pm_synchronize", "pm_stream", "pm_launch_count", "pm_profile_enable",
    "pm_profile_get", "pm_stage_name", "pm_eval_set_coeffs", "pm_eval", "pm_debug_fetch",
    "pm_microbench",
    "pm_timer_start", "pm_timer_stop", "pm_comm_unique_id", "pm_comm_init_rank", "pm_comm_size", "pm_comm_rank", "pm_fit_reduce", "pm_fit_reduce_bytes",
    "pm_comm_allreduce", "pm_comm_barrier", "pm_multi_create", "pm_multi_destroy", "pm_multi_size", "pm_multi_context",
    "pm_multi_fit_reset", "pm_multi_fit_accumulate", "pm_multi_fit_reduce", "pm_multi_fit_finalize",
]


def lib():
    """Load libpolymlp_b200.so (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m pypolymlp_b200.build` "
                "(pypolymlp_b200 has no CPU fallback)."
            )
        L = C.CDLL(LIB_PATH)
        L.pm_last_error.restype = C.c_char_p
        L.pm_version.restype = C.c_char_p
        L.pm_stage_name.restype = C.c_char_p
        L.pm_batch_rows.restype = C.c_int64
        L.pm_launch_count.restype = C.c_int64
        L.pm_stream.restype = C.c_void_p
        L.pm_context_create.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        L.pm_model_create.argtypes = [C.POINTER(FeatureParamsC), C.POINTER(C.c_void_p)]
        L.pm_fit_reduce_bytes.restype = C.c_int64
        L.pm_multi_context.restype = C.c_void_p
        L.pm_multi_context.argtypes = [C.c_void_p, C.c_int]
        L.pm_multi_create.argtypes = [C.c_void_p, _ip, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]
        for name in ("pm_model_destroy", "pm_context_destroy", "pm_multi_destroy"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        _lib = L
    return _lib


class PolymlpB200Error(RuntimeError):
    pass


def check(status):
    if status == 0:
        return
    msg = lib().pm_last_error().decode()
    if status == 1:
        raise ValueError(msg)
    raise PolymlpB200Error(msg)


def as_d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def pd(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def pi(a):
    return a.ctypes.data_as(_ip) if a is not None else None


class StructureBatch:
    """Flat (C ABI) view of a list of structures; keeps the arrays alive."""

    def __init__(self, axis_array, positions_c_array, types_array, force_flags):
        n = len(axis_array)
        if len(positions_c_array) != n or len(types_array) != n or len(force_flags) != n:
            raise ValueError("axis, positions_c, types and force flags must have one entry per structure")
        self.n_st = n
        self.axis = as_d(np.array([np.asarray(a, dtype=np.float64) for a in axis_array])).reshape(n, 9) if n else np.zeros((0, 9))
        self.n_atoms = as_i([np.asarray(p).shape[1] for p in positions_c_array])
        self.positions = (as_d(np.concatenate([as_d(p).reshape(-1) for p in positions_c_array]))
                          if n else np.zeros(0))
        self.types = as_i(np.concatenate([as_i(t).reshape(-1) for t in types_array])) if n else np.zeros(0, np.int32)
        self.force = as_i([int(bool(f)) for f in force_flags])
        for p, t in zip(positions_c_array, types_array):
            if np.asarray(p).shape[0] != 3 or np.asarray(p).shape[1] != len(t):
                raise ValueError("positions_c must be (3, N) with N == len(types)")
        self.c = StructuresC(n, pd(self.axis), pd(self.positions), pi(self.types), pi(self.n_atoms), pi(self.force))

    @property
    def n_rows(self):
        return int(lib().pm_batch_rows(C.byref(self.c)))

    @property
    def h2d_bytes(self):
        return self.axis.nbytes + self.positions.nbytes + self.types.nbytes + self.n_atoms.nbytes + self.force.nbytes
