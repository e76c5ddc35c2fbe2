"""ctypes binding of the C-ABI shared library (include/resnmtf_b200.h).

The library is built in-tree by ``resnmtf_b200/csrc/build.sh`` (``__graft_entry__.build()``).  There is
no CPU fallback anywhere in this package: if the library is missing, or no CUDA device is visible, the
compute entry points raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RESNMTF_B200_LIB") or os.path.join(_HERE, "libresnmtf_b200.so")  # env override: A/B builds of the kernels

OK = 0
E_INVALID, E_CUDA, E_NOMEM, E_STATE, E_NAN, E_UNSUPPORTED, E_COMM = -1, -2, -3, -4, -5, -6, -7
MAX_K = 16
MAP_ROW, MAP_COL = 0, 1
ERR_AUTO, ERR_ALGEBRAIC, ERR_DIRECT = 0, 1, 2
IMPL_AUTO, IMPL_DFMA, IMPL_DMMA, IMPL_TMA, IMPL_FUSED, IMPL_SMALL = 0, 1, 2, 3, 4, 5

# every symbol include/resnmtf_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "resnmtf_ctx_create", "resnmtf_ctx_destroy", "resnmtf_ctx_stream", "resnmtf_ctx_synchronize",
    "resnmtf_ctx_device",
    "resnmtf_last_error", "resnmtf_version", "resnmtf_device_count",
    "resnmtf_fit_create", "resnmtf_fit_create_placed", "resnmtf_fit_destroy", "resnmtf_fit_set_data", "resnmtf_fit_set_data_device",
    "resnmtf_data_create", "resnmtf_data_create_device", "resnmtf_data_destroy", "resnmtf_fit_attach_data",
    "resnmtf_jsd_pairs",
    "resnmtf_data_create_prepped", "resnmtf_data_shape", "resnmtf_data_download", "resnmtf_data_sums",
    "resnmtf_data_shuffle", "resnmtf_data_subsample", "resnmtf_data_copy", "resnmtf_data_svd_topk",
    "resnmtf_data_bisil", "resnmtf_data_bisil_part",
    "resnmtf_pool_create", "resnmtf_pool_destroy", "resnmtf_pool_size", "resnmtf_pool_ctx", "resnmtf_pool_put",
    "resnmtf_pool_put_host", "resnmtf_pool_get", "resnmtf_pool_drop", "resnmtf_batch_run", "resnmtf_unit_size",
    "resnmtf_fit_set_factors", "resnmtf_fit_set_restrictions", "resnmtf_fit_set_shared_map",
    "resnmtf_fit_set_options", "resnmtf_fit_run", "resnmtf_fit_step", "resnmtf_fit_get_factors",
    "resnmtf_fit_normalise", "resnmtf_fit_get_errors", "resnmtf_fit_get_view_errors",
    "resnmtf_fit_get_counters", "resnmtf_fit_profile",
    "resnmtf_comm_id_size", "resnmtf_comm_id_create", "resnmtf_ctx_join",
)


class Counters(C.Structure):
    _fields_ = [
        ("iterations", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("device_ms", C.c_double),
        ("alg_bytes_per_iter", C.c_double),
        ("direct_error_passes", C.c_int64),
        ("converged", C.c_int32),
        ("impl", C.c_int32),
    ]


DERIVE_NONE, DERIVE_SUBSAMPLE, DERIVE_SHUFFLE = 0, 1, 2
DISTANCES = {"euclidean": 0, "manhattan": 1, "cosine": 2}


class Map(C.Structure):
    """resnmtf_map"""
    _fields_ = [("kind", C.c_int32), ("v", C.c_int32), ("w", C.c_int32), ("idx_v", C.c_void_p), ("idx_w", C.c_void_p),
                ("len", C.c_int64)]


class Unit(C.Structure):
    """resnmtf_unit (include/resnmtf_b200.h)"""
    _fields_ = [
        ("data_key", C.c_int32), ("derive", C.c_int32),
        ("rows", C.c_void_p), ("n_rows", C.c_void_p), ("cols", C.c_void_p), ("n_cols", C.c_void_p),
        ("seed", C.c_uint64), ("renormalise", C.c_int32), ("n_maps", C.c_int32),
        ("k", C.c_void_p), ("init_f", C.c_void_p), ("init_s", C.c_void_p), ("init_g", C.c_void_p), ("noise", C.c_void_p),
        ("phi", C.c_void_p), ("xi", C.c_void_p), ("psi", C.c_void_p), ("maps", C.c_void_p),
        ("n_iters", C.c_int64), ("tol", C.c_double), ("max_iters", C.c_int64),
        ("err_mode", C.c_int32), ("impl", C.c_int32),
        ("out_f", C.c_void_p), ("out_s", C.c_void_p), ("out_g", C.c_void_p), ("out_lambda", C.c_void_p),
        ("out_mu", C.c_void_p), ("errors", C.c_void_p), ("errors_cap", C.c_int64), ("n_errors", C.c_int64),
        ("iters", C.c_int64), ("status", C.c_int32), ("gpu", C.c_int32), ("seconds", C.c_double),
        ("message", C.c_char * 200),
    ]


class ResnmtfError(RuntimeError):
    """A C-ABI call returned a negative RESNMTF_E_* code."""

    def __init__(self, code, message):
        super().__init__(f"resnmtf_b200 error {code}: {message}")
        self.code = code
        self.message = message


class ResnmtfNaNError(ResnmtfError, FloatingPointError):
    """The mean error became NaN in convergence mode (the reference's while(NA) at R/main.r:55)."""


_lib = None


def load():
    """Loads the shared library (once).  Raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with resnmtf_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  resnmtf_b200 has no CPU fallback.")
    if not os.environ.get("RESNMTF_NCCL_LIB"):
        # The row-sharded path resolves NCCL with dlopen on first use.  When this interpreter carries a pip-bundled NCCL
        # (the one torch links against), that copy must be the one in the process: an older system libnccl.so.2 loaded
        # first would be handed to a later `import torch` by SONAME and break it.
        import importlib.util

        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec is not None and spec.submodule_search_locations else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["RESNMTF_NCCL_LIB"] = cand
                break
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    pd, pi32, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "resnmtf_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "resnmtf_ctx_destroy": (C.c_int, [vp]),
        "resnmtf_ctx_stream": (vp, [vp]),
        "resnmtf_ctx_synchronize": (C.c_int, [vp]),
        "resnmtf_ctx_device": (C.c_int, [vp]),
        "resnmtf_last_error": (C.c_char_p, []),
        "resnmtf_version": (C.c_char_p, []),
        "resnmtf_device_count": (C.c_int, []),
        "resnmtf_fit_create": (C.c_int, [vp, C.c_int, pi64, pi64, pi32, C.POINTER(vp)]),
        "resnmtf_fit_create_placed": (C.c_int, [vp, C.c_int, pi64, pi64, pi32, C.POINTER(vp)]),
        "resnmtf_fit_destroy": (C.c_int, [vp]),
        "resnmtf_fit_set_data": (C.c_int, [vp, C.c_int, vp, i64]),
        "resnmtf_fit_set_data_device": (C.c_int, [vp, C.c_int, vp, i64]),
        "resnmtf_data_create": (C.c_int, [vp, i64, i64, vp, i64, C.POINTER(vp)]),
        "resnmtf_data_create_device": (C.c_int, [vp, i64, i64, vp, i64, C.POINTER(vp)]),
        "resnmtf_data_destroy": (C.c_int, [vp]),
        "resnmtf_jsd_pairs": (C.c_int, [vp, vp, i64, i32, i64, vp, vp, vp, vp, i64, vp]),
        "resnmtf_fit_attach_data": (C.c_int, [vp, C.c_int, vp]),
        "resnmtf_data_create_prepped": (C.c_int, [vp, i64, i64, vp, i64, pi32, C.POINTER(vp)]),
        "resnmtf_data_shape": (C.c_int, [vp, pi64, pi64]),
        "resnmtf_data_download": (C.c_int, [vp, vp, i64]),
        "resnmtf_data_sums": (C.c_int, [vp, vp, vp]),
        "resnmtf_data_shuffle": (C.c_int, [vp, C.c_uint64, C.c_int, pi64, C.POINTER(vp)]),
        "resnmtf_data_subsample": (C.c_int, [vp, vp, i64, vp, i64, C.POINTER(vp)]),
        "resnmtf_data_copy": (C.c_int, [vp, vp, C.POINTER(vp)]),
        "resnmtf_data_svd_topk": (C.c_int, [vp, C.c_int, vp, vp, vp]),
        "resnmtf_data_bisil": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, pd]),
        "resnmtf_data_bisil_part": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, pi32]),
        "resnmtf_pool_create": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
        "resnmtf_pool_destroy": (C.c_int, [vp]),
        "resnmtf_pool_size": (C.c_int, [vp]),
        "resnmtf_pool_ctx": (vp, [vp, C.c_int]),
        "resnmtf_pool_put": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "resnmtf_pool_put_host": (C.c_int, [vp, C.c_int, C.c_int, pi64, pi64, vp, pi64, C.c_int, pi32]),
        "resnmtf_pool_get": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
        "resnmtf_pool_drop": (C.c_int, [vp, C.c_int]),
        "resnmtf_batch_run": (C.c_int, [vp, vp, C.c_int]),
        "resnmtf_unit_size": (C.c_int, []),
        "resnmtf_fit_set_factors": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
        "resnmtf_fit_set_restrictions": (C.c_int, [vp, vp, vp, vp]),
        "resnmtf_fit_set_shared_map": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, i64]),
        "resnmtf_fit_set_options": (C.c_int, [vp, C.c_int, C.c_int]),
        "resnmtf_fit_run": (C.c_int, [vp, i64, dbl, i64, pi64]),
        "resnmtf_fit_step": (C.c_int, [vp]),
        "resnmtf_fit_get_factors": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
        "resnmtf_fit_normalise": (C.c_int, [vp]),
        "resnmtf_fit_get_errors": (C.c_int, [vp, vp, i64, pi64]),
        "resnmtf_fit_get_view_errors": (C.c_int, [vp, vp, vp]),
        "resnmtf_fit_get_counters": (C.c_int, [vp, C.POINTER(Counters)]),
        "resnmtf_fit_profile": (C.c_int, [vp, i64, pd, pi64]),
        "resnmtf_comm_id_size": (C.c_int, []),
        "resnmtf_comm_id_create": (C.c_int, [vp]),
        "resnmtf_ctx_join": (C.c_int, [vp, vp, C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.resnmtf_unit_size() != C.sizeof(Unit):
        raise RuntimeError("resnmtf_b200: the ctypes declaration of resnmtf_unit does not match the library's")
    _lib = lib
    return lib


def check(code):
    """Turns a negative return code into the matching Python exception."""
    if code == OK:
        return
    msg = load().resnmtf_last_error().decode("utf-8", "replace")
    if code == E_NAN:
        raise ResnmtfNaNError(code, msg)
    raise ResnmtfError(code, msg)


def device_count():
    return int(load().resnmtf_device_count())


_device_checked = False


def require_device():
    """Fails loudly when the CUDA path cannot run (no library, no GPU).  The positive answer is cached: the driver
    query behind it costs ~20 ms and every fit / data handle passes through here."""
    global _device_checked
    lib = load()
    if not _device_checked:
        if lib.resnmtf_device_count() < 1:
            raise RuntimeError("resnmtf_b200: no CUDA device visible and there is no CPU fallback")
        _device_checked = True
    return lib
