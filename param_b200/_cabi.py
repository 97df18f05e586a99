"""ctypes binding of libparam_b200.so (include/param_b200.h).

There is no fallback: if the shared library is missing or a call returns non-zero, this module
raises.  torch is imported first so that the library binds to the CUDA runtime torch already
loaded (one runtime instance => shared streams / current device).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch  # noqa: F401  (must precede the dlopen below)

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libparam_b200.so"

# every symbol include/param_b200.h declares; tests check that the .so exports all of them
EXPORTED_SYMBOLS = (
    "pb200_abi_version", "pb200_error_string", "pb200_launch_count", "pb200_device_info",
    "pb200_embbag_fwd", "pb200_tbe_fwd", "pb200_tbe_fwd_f16", "pb200_check_indices",
    "pb200_tbe_bwd_scratch_bytes", "pb200_tbe_bwd", "pb200_embbag_bwd_sparse", "pb200_tbe_bwd_tables",
    "pb200_tbe_bwd_fused_scratch_bytes", "pb200_tbe_bwd_fused", "pb200_tbe_plan_build",
    "pb200_sort_plan_geometry",
    "pb200_a2a_comm_create", "pb200_a2a_comm_destroy", "pb200_a2a_comm_config",
    "pb200_a2a_comm_error", "pb200_a2a_single", "pb200_a2a_list",
    "pb200_a2a_pooled_fwd", "pb200_a2a_pooled_bwd", "pb200_a2a_pooled_bwd_part", "pb200_tbe_fwd_a2a",
    "pb200_regroup_scratch_bytes", "pb200_regroup_sparse", "pb200_sparse_data_dist",
    "pb200_host_ctx_create", "pb200_host_ctx_destroy", "pb200_tbe_fwd_host", "pb200_tbe_step_host",
    "pb200_tbe_step_host_loss", "pb200_pooled_sum_scratch_bytes", "pb200_pooled_sum",
    "pb200_fill_uniform", "pb200_fill_zipf_indices",
)

POOL_SUM, POOL_MEAN = 0, 1
IDX_I64, IDX_I32 = 0, 1
FWD_AUTO, FWD_DIRECT, FWD_STAGED, FWD_PIPELINED = 0, 1, 2, 3
BWD_AUTO, BWD_ATOMIC, BWD_SORTED, BWD_EXACT = 0, 1, 2, 3
W_F32, W_F16 = 0, 1
OPT_SGD, OPT_ROWWISE_ADAGRAD = 1, 2
A2A_SIGNAL_BYTES = 4096
A2A_MAX_RANKS = 16


class PB200Error(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """dlopen the library (once).  Raises PB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        # a fresh checkout: compile the CUDA sources in-tree if a toolchain is present
        # (this builds the same sm_100a library; it is not a fallback to another code path)
        try:
            from . import build as _build
            _build.build_cuda()
        except Exception as exc:  # noqa: BLE001
            raise PB200Error(
                f"{_LIB_PATH} not found and building it failed ({exc}); build it with "
                "`python -m param_b200.build` (there is no CPU or PyTorch fallback for the "
                "param_b200 kernels)") from exc
    if not _LIB_PATH.exists():
        raise PB200Error(
            f"{_LIB_PATH} not found: build it with `python -m param_b200.build` "
            "(there is no CPU or PyTorch fallback for the param_b200 kernels)"
        )
    lib = C.CDLL(str(_LIB_PATH), mode=os.RTLD_LOCAL | os.RTLD_NOW)
    i32, i64, f32, vp, u64 = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_uint64
    p_i64 = C.POINTER(C.c_int64)

    def sig(name, res, *args):
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("pb200_abi_version", C.c_int)
    sig("pb200_error_string", C.c_char_p, C.c_int)
    sig("pb200_launch_count", i64)
    sig("pb200_device_info", C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
        C.POINTER(C.c_int))
    sig("pb200_embbag_fwd", C.c_int, vp, i64, i32, vp, i64, vp, i64, i32, i32, vp, i32, vp, i64,
        i32, vp)
    sig("pb200_tbe_fwd", C.c_int, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp, i64, i64,
        i32, vp)
    sig("pb200_tbe_fwd_f16", C.c_int, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp, i64, i64,
        vp)
    sig("pb200_check_indices", C.c_int, vp, i32, vp, i64, vp, i64, i32, vp, vp)
    sig("pb200_tbe_bwd_scratch_bytes", i64, i64, i32, i64, i64, i32)
    sig("pb200_tbe_bwd", C.c_int, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp, i64, i64,
        f32, i32, i64, vp, i64, i32, vp)
    sig("pb200_tbe_bwd_tables", C.c_int, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp, i64, i64, f32, i32, i32,
        vp, i64, vp)
    sig("pb200_embbag_bwd_sparse", C.c_int, vp, i64, i32, vp, i64, i32, i64, i32, vp, i32, vp, vp)
    sig("pb200_tbe_bwd_fused_scratch_bytes", i64, i64, i32, i64, i32)
    sig("pb200_tbe_bwd_fused", C.c_int, vp, i32, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, vp,
        i64, i64, i32, f32, f32, i32, u64, i64, vp, i64, i32, vp)
    sig("pb200_tbe_plan_build", C.c_int, vp, i64, vp, i64, i32, i32, vp, i64, vp, i64, i32, vp, i32,
        i64, i64, vp)
    sig("pb200_a2a_comm_create", C.c_int, C.POINTER(vp), i32, i32, C.POINTER(vp), C.POINTER(vp), i64)
    sig("pb200_a2a_comm_destroy", C.c_int, vp)
    sig("pb200_a2a_comm_config", C.c_int, vp, i32, C.c_double)
    sig("pb200_a2a_comm_error", C.c_int, vp, C.POINTER(i32))
    sig("pb200_a2a_single", C.c_int, vp, vp, i64, p_i64, p_i64, i64, vp, vp)
    sig("pb200_a2a_list", C.c_int, vp, C.POINTER(vp), p_i64, p_i64, p_i64, C.POINTER(vp), vp)
    sig("pb200_a2a_pooled_fwd", C.c_int, vp, vp, i64, i64, i32, p_i64, p_i64, i64, vp)
    sig("pb200_a2a_pooled_bwd", C.c_int, vp, vp, i32, p_i64, p_i64, i64, vp)
    sig("pb200_a2a_pooled_bwd_part", C.c_int, vp, vp, i32, p_i64, p_i64, i64, i32, i32, vp)
    sig("pb200_tbe_fwd_a2a", C.c_int, vp, vp, vp, i32, i32, vp, i64, vp, i32, i32, p_i64, p_i64, i64, vp)
    sig("pb200_regroup_scratch_bytes", i64, i32, i32, i64)
    sig("pb200_regroup_sparse", C.c_int, vp, vp, i64, i32, i32, i64, vp, vp, vp, vp, i64, vp)
    sig("pb200_sparse_data_dist", C.c_int, vp, vp, vp, i64, p_i64, i64, i64, i64, i64, vp, vp, vp, vp, i64,
        vp)
    sig("pb200_host_ctx_create", C.c_int, C.POINTER(vp), i64, i64, i32)
    sig("pb200_host_ctx_destroy", C.c_int, vp)
    sig("pb200_tbe_fwd_host", C.c_int, vp, vp, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, i32)
    sig("pb200_tbe_step_host", C.c_int, vp, vp, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, i32,
        i32, f32)
    sig("pb200_tbe_step_host_loss", C.c_int, vp, vp, vp, vp, i32, i32, vp, i64, vp, i64, i32, vp, i32, i32, f32)
    p_i32 = C.POINTER(C.c_int32)
    sig("pb200_sort_plan_geometry", C.c_int, i64, i32, i64, i64, p_i32, p_i32, p_i32, p_i32, p_i32, p_i64, p_i64)
    sig("pb200_pooled_sum_scratch_bytes", i64, i64)
    sig("pb200_pooled_sum", C.c_int, vp, i64, i64, vp, vp, i64, vp)
    sig("pb200_fill_uniform", C.c_int, vp, i64, f32, f32, u64, vp)
    sig("pb200_fill_zipf_indices", C.c_int, vp, i64, i32, vp, i64, i32, u64, vp)
    if lib.pb200_abi_version() != 2:
        raise PB200Error("libparam_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().pb200_error_string(rc).decode()
        raise PB200Error(f"{what or 'pb200 call'} failed: {msg} (code {rc})")


def launch_count() -> int:
    return int(load().pb200_launch_count())


def i64_array(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr
