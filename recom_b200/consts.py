"""Enums and record layouts of include/recom_b200.h as plain Python (no native code).

Split from `_lib` so that the workload generators (`schema`, `synth`) can be used by the CPU
reference arm of bench.py without mapping librecom_b200.so into that process. The product
package itself (`import recom_b200`) always loads the library and fails loudly without it.
"""
from __future__ import annotations

import numpy as np

# enums (include/recom_b200.h)
OK = 0
ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_WORKSPACE_TOO_SMALL = 1, 2, 3
ERR_UNSUPPORTED, ERR_OUT_OF_MEMORY, ERR_ID_OUT_OF_RANGE = 4, 5, 6
F32, BF16, I32, I64, STR, F64 = 0, 1, 2, 3, 4, 5
LOOKUP_DENSE, LOOKUP_GATHER_SCATTER, LOOKUP_SEGMENT = 0, 1, 2
COMBINE_SUM, COMBINE_MEAN, COMBINE_SQRTN = 0, 1, 2
XFORM_NONE, XFORM_BUCKETIZE, XFORM_HASH_STR, XFORM_HASH_INT = 0, 1, 2, 3
SEG_NONE, SEG_ROW_SPLITS_I32, SEG_COO_I64, SEG_SPLIT_STR = 0, 1, 2, 3
SEG_COO_RESHAPE_I64, SEG_FIXED = 4, 5


def reshape_spec(rank: int, has_new_shape: bool = False) -> int:
    """RECOM_RESHAPE_SPEC (include/recom_b200.h)."""
    return rank | (256 if has_new_shape else 0)

DTYPE_BYTES = {F32: 4, BF16: 2, I32: 4, I64: 8, STR: 4, F64: 8}

# numpy mirror of struct recom_column_desc (natural C alignment)
COLUMN_DESC = np.dtype(
    [
        ("lookup", "<i4"), ("combiner", "<i4"), ("transform", "<i4"), ("value_dtype", "<i4"),
        ("table", "<i4"), ("table_dtype", "<i4"), ("table_rows", "<i8"),
        ("dim", "<i4"), ("out_offset", "<i4"), ("value_slot", "<i4"), ("bytes_slot", "<i4"),
        ("seg_slot", "<i4"), ("seg_format", "<i4"), ("seg_stride", "<i4"),
        ("num_boundaries", "<i4"), ("boundaries_offset", "<i8"), ("num_buckets", "<i8"),
    ],
    align=True,
)
assert COLUMN_DESC.itemsize == 80



# pre-lookup id filters (recom_column_filter)
FILTER_NONE, FILTER_KEEP, FILTER_SELECT, FILTER_STR_NE = 0, 1, 2, 3
MAX_INTERVALS = 4
COLUMN_FILTER = np.dtype(
    [
        ("kind", "<i4"), ("num_intervals", "<i4"), ("lo", "<i8", (MAX_INTERVALS,)),
        ("hi", "<i8", (MAX_INTERVALS,)), ("substitute", "<i8"), ("not_equal_len", "<i4"),
        ("not_equal", "u1", (16,)), ("pad", "<i4"),
    ],
    align=True,
)
assert COLUMN_FILTER.itemsize == 104


def split_spec(delimiter: bytes, skip_empty: bool = True) -> int:
    """RECOM_SPLIT_SPEC of include/recom_b200.h: up to 3 delimiter characters."""
    d = bytes(delimiter)
    assert len(d) <= 3
    d3 = d + b"\0" * (3 - len(d))
    return d3[0] | d3[1] << 8 | d3[2] << 16 | len(d) << 24 | (1 if skip_empty else 0) << 28
