"""ctypes binding of ``libged_b200.so`` (the C-ABI declared in ``include/ged_b200.h``).

There is deliberately no CPU fallback: if the CUDA library is missing or no CUDA device is present every compute
entry point raises.  The library is built in-tree by ``build.py`` (``nvcc -gencode arch=compute_100a,code=sm_100a``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GED_B200_LIBRARY") or os.path.join(_HERE, "libged_b200.so")     # override: A/B of kernel builds

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_u64p = ctypes.POINTER(ctypes.c_ulonglong)

GED_OK, GED_E_BADARG, GED_E_TOO_LARGE, GED_E_CUDA = 0, 1, 2, 3
GED_ST_INVALID_COST, GED_ST_INFEASIBLE, GED_ST_BAD_GRAPH = 1, 2, 4
GED_SOLVER_IPFP, GED_SOLVER_HUNGARIAN = 0, 1
ABI_VERSION = 9

# every symbol include/ged_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = (
    "ged_solve_batch", "ged_lsape_batch", "ged_kv_batch", "ged_sinkhorn_batch", "ged_score_batch", "ged_quadform_ws_elems",
    "ged_quadform_dense_batch", "ged_limits", "ged_solve_smem_bytes", "ged_solve_max_resident", "ged_abi_version", "ged_version",
    "ged_last_error",
)


class GedPairs(ctypes.Structure):
    _fields_ = [
        ("num_pairs", ctypes.c_int32),
        ("n1", c_i32p), ("n2", c_i32p),
        ("adj1", c_u8p), ("adj1_off", c_i64p),
        ("adj2", c_u8p), ("adj2_off", c_i64p),
        ("lab1", c_i32p), ("lab1_off", c_i64p),
        ("lab2", c_i32p), ("lab2_off", c_i64p),
        ("node_cost", c_f32p), ("node_cost_off", c_i64p),
        ("max_n1", ctypes.c_int32), ("max_n2", ctypes.c_int32),
        ("adj1_ld", ctypes.c_int32), ("adj2_ld", ctypes.c_int32),
    ]


class GedSolveDesc(ctypes.Structure):
    _fields_ = [
        ("pairs", GedPairs),
        ("B", ctypes.c_int32),
        ("base_idx", c_i32p),
        ("edit_u", c_i32p), ("edit_v", c_i32p),
        ("score_adj1", c_u8p), ("score_adj1_off", c_i64p), ("score_idx", c_i32p),
        ("score_adj1_ld", ctypes.c_int32),
        ("max_iter", ctypes.c_int32),
        ("solver_kind", ctypes.c_int32),
        ("x_init", c_f32p),
        ("mat_stride", ctypes.c_int64),
        ("workspace", c_f32p),
        ("workspace_locks", c_i32p),
        ("workspace_slots", ctypes.c_int32),
        ("workspace_stride", ctypes.c_int64),
        ("order", c_i32p),
        ("order_count", ctypes.c_int32),
        ("work_counter", c_i32p),
        ("launch_hint", ctypes.c_int32),
        ("workspace_mats", ctypes.c_int32),
    ]


class GedSolveOut(ctypes.Structure):
    _fields_ = [
        ("ged", c_f32p), ("ged_edited", c_f32p),
        ("rowmatch", c_i32p), ("colmatch", c_i32p),
        ("match_stride_1", ctypes.c_int32), ("match_stride_2", ctypes.c_int32),
        ("iters", c_i32p), ("status", c_i32p),
        ("x_out", c_f32p),
        ("tr_cost", c_f32p), ("tr_x_after", c_f32p), ("tr_scalars", c_f64p),
        ("tr_rowmatch", c_i32p), ("tr_colmatch", c_i32p),
        ("tr_init_cost", c_f32p), ("tr_init_rowmatch", c_i32p), ("tr_init_colmatch", c_i32p),
        ("tr_lap_steps", c_u64p),
    ]


_lib = None


class GedLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (no GPU needed for this; used by the CPU symbol test)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GedLibraryError(
            f"{LIB_PATH} is missing: build it with `python ppo-bihyb_b200/build.py` (or __graft_entry__.build()). "
            "There is no CPU fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.ged_solve_batch.argtypes = [ctypes.POINTER(GedSolveDesc), ctypes.POINTER(GedSolveOut), ctypes.c_void_p]
    lib.ged_solve_batch.restype = ctypes.c_int
    lib.ged_lsape_batch.argtypes = [ctypes.c_int32, c_i32p, c_i32p, ctypes.c_int32, ctypes.c_int32, c_f32p,
                                    ctypes.c_int64, c_i32p, ctypes.c_int32, c_i32p, ctypes.c_int32, c_f32p, c_i32p,
                                    ctypes.c_void_p]
    lib.ged_lsape_batch.restype = ctypes.c_int
    lib.ged_kv_batch.argtypes = [ctypes.POINTER(GedPairs), c_f32p, c_f32p, c_f32p, ctypes.c_int64, ctypes.c_void_p]
    lib.ged_kv_batch.restype = ctypes.c_int
    lib.ged_sinkhorn_batch.argtypes = [ctypes.c_int32, c_i32p, c_i32p, ctypes.c_int32, ctypes.c_int32, c_f32p, c_f32p, ctypes.c_int32,
                                       c_f32p, ctypes.c_void_p]
    lib.ged_sinkhorn_batch.restype = ctypes.c_int
    lib.ged_score_batch.argtypes = [ctypes.POINTER(GedPairs), ctypes.c_int32, c_i32p, c_i32p, ctypes.c_int32,
                                    c_i32p, ctypes.c_int32, c_f32p, ctypes.c_void_p]
    lib.ged_score_batch.restype = ctypes.c_int
    lib.ged_quadform_ws_elems.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.ged_quadform_ws_elems.restype = ctypes.c_int64
    lib.ged_quadform_dense_batch.argtypes = [ctypes.c_int32, ctypes.c_int32, c_f32p, ctypes.c_int64, c_f32p, c_f32p,
                                             c_f32p, ctypes.c_void_p]
    lib.ged_quadform_dense_batch.restype = ctypes.c_int
    lib.ged_limits.argtypes = [c_i32p, c_i32p]
    lib.ged_limits.restype = None
    lib.ged_solve_max_resident.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.ged_solve_max_resident.restype = ctypes.c_int32
    lib.ged_solve_smem_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    lib.ged_solve_smem_bytes.restype = ctypes.c_int64
    lib.ged_abi_version.restype = ctypes.c_int
    lib.ged_version.restype = ctypes.c_char_p
    lib.ged_last_error.restype = ctypes.c_char_p
    if lib.ged_abi_version() != ABI_VERSION:
        raise GedLibraryError(f"ABI mismatch: library {lib.ged_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def require_cuda():
    """The compute path: library present AND a CUDA device.  Raises otherwise (no fallback)."""
    import torch
    lib = load()
    if not torch.cuda.is_available():
        raise GedLibraryError("no CUDA device: the GED solver path has no CPU fallback")
    return lib


def check(rc: int, what: str):
    if rc == GED_OK:
        return
    msg = load().ged_last_error().decode()
    if rc == GED_E_BADARG:
        raise ValueError(f"{what}: {msg}")
    if rc == GED_E_TOO_LARGE:
        raise NotImplementedError(f"{what}: {msg}")
    raise GedLibraryError(f"{what}: {msg}")


def ptr(t, ctype):
    """Device pointer of a torch tensor (or None) as a typed ctypes pointer."""
    if t is None:
        return ctypes.cast(None, ctype)
    return ctypes.cast(ctypes.c_void_p(t.data_ptr()), ctype)
