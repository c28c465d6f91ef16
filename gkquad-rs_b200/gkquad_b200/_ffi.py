"""ctypes binding of libgkquad_b200.so (the C ABI declared in include/gkquad_b200.h).

There is NO fallback: if the CUDA library has not been built, or no CUDA device is present,
importing works (so that host-side logic can be tested) but every compute entry raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
import os

# GKQ_LIB selects an alternative build of the same library (kernel-tuning variants)
LIB_PATH = Path(os.environ["GKQ_LIB"]) if os.environ.get("GKQ_LIB") else _PKG.parent / "lib" / "libgkquad_b200.so"

GKQ_OK = 0
ERRORS = {-1: "GKQ_ERR_INVALID_ARGUMENT", -2: "GKQ_ERR_NAN_INPUT", -3: "GKQ_ERR_NO_DEVICE",
          -4: "GKQ_ERR_CUDA", -5: "GKQ_ERR_OUT_OF_MEMORY", -6: "GKQ_ERR_UNSUPPORTED"}


class GkqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class gkq_tolerance(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("abs_tol", C.c_double), ("rel_tol", C.c_double)]


class gkq_config(C.Structure):
    _fields_ = [("algo", C.c_int32), ("flags", C.c_int32), ("tol", gkq_tolerance),
                ("max_evals", C.c_uint64), ("workspace_capacity", C.c_uint64),
                ("points", C.POINTER(C.c_double)), ("npoints", C.c_int64)]


class gkq_overrides(C.Structure):
    """Per-integral tolerance values / budgets (NULL = the batch value); include/gkquad_b200.h."""
    _fields_ = [("abs_tol", C.c_void_p), ("rel_tol", C.c_void_p), ("max_evals", C.c_void_p)]


class gkq_stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("batches", C.c_uint64), ("integrals", C.c_uint64),
                ("scratch_bytes", C.c_uint64), ("last_survivors_qk25", C.c_uint64),
                ("last_survivors_loop", C.c_uint64), ("last_survivors_step2", C.c_uint64),
                ("last_chain_loop", C.c_uint64 * 2), ("last_chain_step2", C.c_uint64 * 2)]


# every symbol include/gkquad_b200.h declares (tests/test_abi.py checks the list against the header)
SYMBOLS = [
    "gkq_abi_version", "gkq_config_default", "gkq_create", "gkq_create_prio", "gkq_destroy", "gkq_last_error",
    "gkq_functor_count", "gkq_functor_info", "gkq_functor_lookup", "gkq_integrate_batch",
    "gkq_integrate_batch_host", "gkq_integrate_batch_ex", "gkq_integrate_batch_host_ex", "gkq_integrate2_batch_ex", "gkq_integrate2_batch_host_ex", "gkq_integrate2_batch", "gkq_integrate2_batch_host", "gkq_qk_batch",
    "gkq_functor_eval_batch", "gkq_functor_eval_batch_ex",
    "gkq_set_deferred_join", "gkq_join", "gkq_sync", "gkq_get_stats", "gkq_reset_stats", "gkq_set_profiling", "gkq_get_profile",
    "gkq_measure_fp64_peak",
    "gkq_comm_get_unique_id", "gkq_comm_init", "gkq_comm_destroy", "gkq_comm_info", "gkq_shard_size",
    "gkq_shard_index", "gkq_gather", "gkq_gather_range",
]
GKQ_COMM_ID_BYTES = 128

_lib = None


def lib():
    """Load the product library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first (make -C gkquad-rs_b200, or "
            "__graft_entry__.build()).  gkquad_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, dp, u32p, i64p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    L.gkq_abi_version.restype = C.c_int
    L.gkq_config_default.argtypes = [C.c_int, C.POINTER(gkq_config)]
    L.gkq_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.gkq_create_prio.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    L.gkq_destroy.argtypes = [vp]
    L.gkq_last_error.argtypes = [vp]
    L.gkq_last_error.restype = C.c_char_p
    L.gkq_functor_count.argtypes = [C.c_int]
    L.gkq_functor_info.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                   C.POINTER(C.c_double)]
    L.gkq_functor_lookup.argtypes = [C.c_int, C.c_char_p]
    batch1 = [vp, C.POINTER(gkq_config), C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64, i64p,
              dp, dp, dp, u32p, u32p]
    L.gkq_integrate_batch.argtypes = batch1 + [vp]
    L.gkq_integrate_batch_host.argtypes = batch1
    batch1x = batch1[:2] + [C.POINTER(gkq_overrides)] + batch1[2:]
    L.gkq_integrate_batch_ex.argtypes = batch1x + [vp]
    L.gkq_integrate_batch_host_ex.argtypes = batch1x
    batch2 = [vp, C.POINTER(gkq_config), C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64,
              dp, C.c_int64, i64p, dp, dp, dp, u32p, u32p]
    L.gkq_integrate2_batch.argtypes = batch2 + [vp]
    L.gkq_integrate2_batch_host.argtypes = batch2
    batch2x = batch2[:2] + [C.POINTER(gkq_overrides)] + batch2[2:]
    L.gkq_integrate2_batch_ex.argtypes = batch2x + [vp]
    L.gkq_integrate2_batch_host_ex.argtypes = batch2x
    L.gkq_qk_batch.argtypes = [vp, C.c_int, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int64, dp, vp]
    L.gkq_functor_eval_batch.argtypes = [vp, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, vp]
    L.gkq_functor_eval_batch_ex.argtypes = [vp, C.c_int, C.c_int64, dp, dp, C.c_int64, dp, C.c_int, vp]
    L.gkq_set_deferred_join.argtypes = [vp, C.c_int]
    L.gkq_join.argtypes = [vp, vp]
    L.gkq_sync.argtypes = [vp, vp]
    L.gkq_get_stats.argtypes = [vp, C.POINTER(gkq_stats)]
    L.gkq_reset_stats.argtypes = [vp]
    L.gkq_set_profiling.argtypes = [vp, C.c_int]
    L.gkq_get_profile.argtypes = [vp, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.gkq_measure_fp64_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.gkq_comm_get_unique_id.argtypes = [vp]
    L.gkq_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gkq_comm_destroy.argtypes = [vp]
    L.gkq_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gkq_shard_size.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64]
    L.gkq_shard_size.restype = C.c_int64
    L.gkq_shard_index.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64]
    L.gkq_shard_index.restype = C.c_int64
    L.gkq_gather.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, dp, dp, u32p, u32p, dp, dp, u32p, u32p, vp]
    L.gkq_gather_range.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_int64, dp, dp, u32p, u32p,
                                   dp, dp, u32p, u32p, vp]
    _lib = L
    return L


def functor_table(dim: int):
    """[(id, name, nparams, flops_per_eval)]; dim 1 = 1-D, 2 = 2-D, 0 = y-range functors.
    Pure host query: works without a GPU."""
    L = lib()
    out = []
    for i in range(L.gkq_functor_count(dim)):
        name, npar, fl = C.c_char_p(), C.c_int(), C.c_double()
        assert L.gkq_functor_info(dim, i, C.byref(name), C.byref(npar), C.byref(fl)) == 0
        out.append((i, name.value.decode(), npar.value, fl.value))
    return out
