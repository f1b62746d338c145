"""ctypes binding of librecom_b200.so (the C ABI in include/recom_b200.h).

There is no fallback: if the shared library (hand-written sm_100a kernels + C++ host
code) is missing or fails to load, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RECOM_B200_LIB selects another build of the same library (e.g. the phase-timer build)
LIB_PATH = os.environ.get("RECOM_B200_LIB") or os.path.join(_HERE, "librecom_b200.so")

from .consts import *  # noqa: F401,F403  (enums, COLUMN_DESC, DTYPE_BYTES)
from .consts import COLUMN_DESC, DTYPE_BYTES  # noqa: F401


class PlanConfig(C.Structure):
    _fields_ = [
        ("out_dtype", C.c_int32), ("out_width", C.c_int32), ("num_slots", C.c_int32),
        ("num_tables", C.c_int32), ("check_ids", C.c_int32), ("device", C.c_int32),
        ("max_ctas_per_sm", C.c_int32), ("row_shard", C.c_int32),
    ]


class Request(C.Structure):
    _fields_ = [
        ("batch", C.c_int64), ("blob", C.c_void_p), ("slot_offsets", C.c_void_p),
        ("slot_elems", C.c_void_p), ("tables", C.c_void_p), ("out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t), ("ids_out", C.c_void_p),
        ("status_out", C.c_void_p), ("stream", C.c_void_p), ("args_offset", C.c_int64),
    ]


class ShardOut(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("col_offset", C.c_int32), ("global_width", C.c_int32),
        ("rank", C.c_int32), ("rows_per_rank", C.c_int64), ("peer_out", C.c_void_p),
        ("peer_flags", C.c_void_p), ("epoch", C.c_uint32), ("flags", C.c_uint32),
    ]


SHARD_FLAG_WORDS = 16
SHARD_WAIT_IN_KERNEL = 1


class RecomError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"recom_b200 status {status}: {message}")
        self.status = status


EXPORTS = [
    "recom_last_error", "recom_abi_version", "recom_plan_create", "recom_plan_create_filtered", "recom_plan_destroy",
    "recom_plan_serialize", "recom_plan_create_from_blob", "recom_plan_num_columns",
    "recom_plan_out_width", "recom_plan_num_slots", "recom_plan_slot_layout", "recom_plan_infer_batch",
    "recom_plan_output_shapes", "recom_plan_output_ptrs", "recom_plan_scale_rows", "recom_plan_workspace_bytes", "recom_plan_run",
    "recom_last_launch_count", "recom_plan_algorithmic_bytes", "recom_scatter_columns", "recom_scatter_host_columns",
    "recom_segment_ids_to_row_splits", "recom_string_to_hash_bucket_fast", "recom_fingerprint64",
    "recom_concat_inputs", "recom_concat_inputs_with_args", "recom_shape_construct_create", "recom_shape_construct_eval",
    "recom_shape_construct_destroy", "recom_plan_run_sharded", "recom_ipc_alloc", "recom_ipc_open",
    "recom_ipc_close", "recom_ipc_free", "recom_shard_wait", "recom_dense_relu",
    "recom_plan_set_request_overlap",
]


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C recom_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "recom_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.recom_last_error.restype = C.c_char_p
    lib.recom_abi_version.restype = C.c_int
    lib.recom_last_launch_count.restype = C.c_int
    vp, i32, i64, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
    sigs = {
        "recom_plan_create": [vp, i32, vp, i64, C.POINTER(PlanConfig), C.POINTER(vp)],
        "recom_plan_create_filtered": [vp, i32, vp, i64, vp, C.POINTER(PlanConfig), C.POINTER(vp)],
        "recom_plan_destroy": [vp],
        "recom_plan_serialize": [vp, vp, sz, C.POINTER(sz)],
        "recom_plan_create_from_blob": [vp, sz, i32, C.POINTER(vp)],
        "recom_plan_num_columns": [vp, C.POINTER(i32)],
        "recom_plan_out_width": [vp, C.POINTER(i32)],
        "recom_plan_num_slots": [vp, C.POINTER(i32)],
        "recom_plan_slot_layout": [vp, vp, vp],
        "recom_plan_infer_batch": [vp, vp, C.POINTER(i64)],
        "recom_plan_output_shapes": [vp, i64, vp],
        "recom_plan_output_ptrs": [vp, vp, vp, vp],
        "recom_plan_scale_rows": [vp, C.POINTER(Request), i64, i64, vp],
        "recom_plan_workspace_bytes": [vp, i64, i64, C.POINTER(sz)],
        "recom_plan_run": [vp, C.POINTER(Request)],
        "recom_plan_run_sharded": [vp, C.POINTER(Request), C.POINTER(ShardOut)],
        "recom_shard_wait": [vp, i32, C.c_uint32, C.c_uint32, vp],
        "recom_ipc_alloc": [sz, C.POINTER(vp), vp],
        "recom_ipc_open": [vp, C.POINTER(vp)],
        "recom_ipc_close": [vp],
        "recom_ipc_free": [vp],
        "recom_plan_algorithmic_bytes": [vp, C.POINTER(Request), C.POINTER(i64)],
        "recom_scatter_host_columns": [vp, vp, vp, i32, i64, i32, i32, vp, vp, vp],
        "recom_scatter_columns": [vp, vp, vp, i32, i64, i32, i32, vp, vp],
        "recom_dense_relu": [vp, i64, i32, i64, vp, i32, vp, i32, vp, vp],
        "recom_plan_set_request_overlap": [vp, i32],
        "recom_segment_ids_to_row_splits": [vp, i64, i32, i64, vp, vp],
        "recom_string_to_hash_bucket_fast": [vp, vp, i64, i64, vp, vp],
        "recom_fingerprint64": [vp, vp, i64, vp, vp],
        "recom_concat_inputs": [vp, vp, i32, vp, sz, vp, C.POINTER(sz)],
        "recom_concat_inputs_with_args": [vp, vp, vp, i32, vp, sz, vp, C.POINTER(i64), C.POINTER(sz)],
        "recom_shape_construct_create": [vp, i32, vp, i32, C.POINTER(vp)],
        "recom_shape_construct_eval": [vp, vp, i32, vp, i32],
        "recom_shape_construct_destroy": [vp],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    return lib


lib = _load()


def check(status: int) -> None:
    if status != OK:
        raise RecomError(status, lib.recom_last_error().decode("utf-8", "replace"))
