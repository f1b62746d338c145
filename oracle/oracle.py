"""ctypes wrapper of oracle/librecom_oracle.so — TEST INFRASTRUCTURE ONLY.

numpy in, numpy out. Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librecom_oracle.so")

F32, BF16, I32, I64, STR, F64 = 0, 1, 2, 3, 4, 5


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "recom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load() -> C.CDLL:
    build()
    lib = C.CDLL(_SO)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.oracle_fingerprint64.restype = C.c_uint64
    lib.oracle_fingerprint64.argtypes = [C.c_char_p, C.c_size_t]
    lib.oracle_fingerprint64_batch.argtypes = [vp, vp, i64, vp]
    lib.oracle_string_to_hash_bucket_fast.argtypes = [vp, vp, i64, i64, vp]
    lib.oracle_hash_int_i32.argtypes = [vp, i64, i64, vp]
    lib.oracle_hash_int_i64.argtypes = [vp, i64, i64, vp]
    for n in ("f32", "f64", "i32", "i64"):
        getattr(lib, f"oracle_bucketize_{n}").argtypes = [vp, i64, vp, i32, vp]
    lib.oracle_f32_to_bf16.argtypes = [vp, i64, vp]
    lib.oracle_bf16_to_f32.argtypes = [vp, i64, vp]
    lib.oracle_sparse_fill_empty_rows.restype = i64
    lib.oracle_sparse_fill_empty_rows.argtypes = [vp, vp, i64, i32, i64, i64, vp, vp, vp, vp]
    lib.oracle_sparse_segment_reduce.restype = i64
    lib.oracle_sparse_segment_reduce.argtypes = [vp, i32, i64, i32, vp, vp, i64, i64, i32, vp]
    lib.oracle_safe_embedding_lookup_sparse.restype = C.c_int
    lib.oracle_safe_embedding_lookup_sparse.argtypes = [vp, i32, i64, i32, vp, vp, i64, i32, i64, i32, vp]
    lib.oracle_recom_lookup.restype = C.c_int
    lib.oracle_recom_lookup.argtypes = [vp, i32, i64, i32, i32, i32, vp, vp, i64, i64, vp]
    lib.oracle_run_model.restype = C.c_int
    lib.oracle_run_model.argtypes = [vp, i32, vp, i32, i32, i64, vp, vp, vp, vp, vp, vp, i32]
    lib.oracle_num_threads.restype = C.c_int
    return lib


lib = _load()


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def fingerprint64(s: bytes) -> int:
    return int(lib.oracle_fingerprint64(s, len(s)))


def fingerprint64_batch(data: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    n = offsets.size - 1
    out = np.zeros(n, np.uint64)
    data = np.ascontiguousarray(data, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32)
    lib.oracle_fingerprint64_batch(_p(data), _p(offsets), n, _p(out))
    return out


def string_to_hash_bucket_fast(data: np.ndarray, offsets: np.ndarray, num_buckets: int) -> np.ndarray:
    n = offsets.size - 1
    out = np.zeros(n, np.int64)
    data = np.ascontiguousarray(data, np.uint8)
    offsets = np.ascontiguousarray(offsets, np.int32)
    lib.oracle_string_to_hash_bucket_fast(_p(data), _p(offsets), n, num_buckets, _p(out))
    return out


def hash_int(values: np.ndarray, num_buckets: int) -> np.ndarray:
    out = np.zeros(values.size, np.int64)
    if values.dtype == np.int32:
        v = np.ascontiguousarray(values)
        lib.oracle_hash_int_i32(_p(v), v.size, num_buckets, _p(out))
    else:
        v = np.ascontiguousarray(values, np.int64)
        lib.oracle_hash_int_i64(_p(v), v.size, num_buckets, _p(out))
    return out


def bucketize(values: np.ndarray, boundaries: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(boundaries, np.float32)
    v = np.ascontiguousarray(values)
    out = np.zeros(v.size, np.int32)
    name = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64",
            np.dtype(np.int32): "i32", np.dtype(np.int64): "i64"}[v.dtype]
    getattr(lib, f"oracle_bucketize_{name}")(_p(v), v.size, _p(b), b.size, _p(out))
    return out


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.shape, np.uint16)
    lib.oracle_f32_to_bf16(_p(x), x.size, _p(out))
    return out


def bf16_bits_to_f32(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.uint16)
    out = np.zeros(x.shape, np.float32)
    lib.oracle_bf16_to_f32(_p(x), x.size, _p(out))
    return out


def _table_arg(table: np.ndarray) -> Tuple[np.ndarray, int]:
    if table.dtype == np.uint16:
        return np.ascontiguousarray(table), BF16
    return np.ascontiguousarray(table, np.float32), F32


def sparse_fill_empty_rows(indices: np.ndarray, values: np.ndarray, dense_rows: int, default: int = 0):
    indices = np.ascontiguousarray(indices, np.int64).reshape(values.size, -1)
    values = np.ascontiguousarray(values, np.int64)
    rank = indices.shape[1] if indices.ndim == 2 and indices.shape[1] else 2
    cap = values.size + dense_rows
    oi = np.zeros((cap, rank), np.int64)
    ov = np.zeros(cap, np.int64)
    empty = np.zeros(dense_rows, np.uint8)
    rev = np.zeros(values.size, np.int64)
    n = lib.oracle_sparse_fill_empty_rows(_p(indices), _p(values), values.size, rank, dense_rows,
                                          default, _p(oi), _p(ov), _p(empty), _p(rev))
    if n < 0:
        raise ValueError("invalid sparse indices")
    return oi[:n], ov[:n], empty.astype(bool), rev


def sparse_segment_reduce(table: np.ndarray, idx: np.ndarray, segment_ids: np.ndarray,
                          num_segments: int, combiner: int) -> np.ndarray:
    t, dt = _table_arg(table)
    idx = np.ascontiguousarray(idx, np.int64)
    seg = np.ascontiguousarray(segment_ids, np.int64)
    rows = num_segments if num_segments >= 0 else (int(seg[-1]) + 1 if seg.size else 0)
    out = np.zeros((rows, t.shape[1]), np.float32)
    r = lib.oracle_sparse_segment_reduce(_p(t), dt, t.shape[0], t.shape[1], _p(idx), _p(seg),
                                         idx.size, num_segments, combiner, _p(out))
    if r < 0:
        raise ValueError("invalid segment reduction input")
    return out


def safe_embedding_lookup_sparse(table: np.ndarray, indices: np.ndarray, ids: np.ndarray,
                                 dense_rows: int, combiner: int) -> np.ndarray:
    t, dt = _table_arg(table)
    ids = np.ascontiguousarray(ids, np.int64)
    indices = np.ascontiguousarray(indices, np.int64).reshape(ids.size, -1)
    rank = indices.shape[1] if indices.shape[1] else 2
    out = np.zeros((dense_rows, t.shape[1]), np.float32)
    rc = lib.oracle_safe_embedding_lookup_sparse(_p(t), dt, t.shape[0], t.shape[1], _p(indices),
                                                 _p(ids), ids.size, rank, dense_rows, combiner, _p(out))
    if rc != 0:
        raise ValueError("invalid lookup input")
    return out


def recom_lookup(table: np.ndarray, lookup: int, combiner: int, ids: np.ndarray,
                 row_of: Optional[np.ndarray], batch: int) -> np.ndarray:
    t, dt = _table_arg(table)
    ids = np.ascontiguousarray(ids, np.int64)
    ro = None if row_of is None else np.ascontiguousarray(row_of, np.int64)
    out = np.zeros((batch, t.shape[1]), np.float32)
    rc = lib.oracle_recom_lookup(_p(t), dt, t.shape[0], t.shape[1], lookup, combiner, _p(ids),
                                 _p(ro), ids.size, batch, _p(out))
    if rc != 0:
        raise ValueError("invalid lookup input")
    return out


def run_model(descs: np.ndarray, boundaries: np.ndarray, out_width: int, out_dtype: int, batch: int,
              blob: np.ndarray, slot_offsets: np.ndarray, slot_elems: np.ndarray,
              tables: Sequence[np.ndarray], total_ids: int, num_threads: int = 0):
    """Whole stage on the host. tables: float32 arrays, or uint16 (bf16 bits).
    -> (out float32 [batch, out_width], ids int64 [total_ids])."""
    keep = [np.ascontiguousarray(t) for t in tables]
    ptrs = np.array([t.ctypes.data for t in keep], dtype=np.uint64)
    out = np.zeros((batch, out_width), np.float32)
    ids = np.zeros(max(total_ids, 1), np.int64)
    b = np.ascontiguousarray(boundaries, np.float32)
    rc = lib.oracle_run_model(_p(descs), len(descs), _p(b) if b.size else None, out_width, out_dtype,
                              batch, _p(blob), _p(slot_offsets), _p(slot_elems), _p(ptrs), _p(out),
                              _p(ids), num_threads)
    if rc != 0:
        raise ValueError("oracle_run_model failed (invalid input)")
    return out, ids[:total_ids]


def num_threads() -> int:
    return int(lib.oracle_num_threads())


# ---------------------------------------------------------------------------------------
# The reference's pre-lookup filter ops, restated. TEST INFRASTRUCTURE ONLY.
# The comparator of the reference kernels reads `x >= l || x <= r`, which holds for every x; the
# interval sets the graph pass hands them (pre_lookup_optimizer.cc:66-123: GreaterEqual -> [c, MAX],
# LessEqual -> [MIN, c], ...) and the graph they replace (tf.where / sparse.retain) mean
# `l <= x <= r`: that is what is restated here.
# ---------------------------------------------------------------------------------------
def _in_intervals(values: np.ndarray, intervals) -> np.ndarray:
    keep = np.zeros(values.shape, bool)
    for lo, hi in intervals:
        keep |= (values >= lo) & (values <= hi)
    return keep


def gather_indice_value(indices: np.ndarray, values: np.ndarray, intervals):
    """Addons>GatherIndiceValue (custom_ops/gather_indice_value/gather_indice_value_ops.cc:33-70):
    (indices [nnz, rank], values [nnz]) -> the entries whose value lies in the interval set."""
    keep = _in_intervals(values, intervals)
    return indices[keep], values[keep]


def select_value(values: np.ndarray, intervals, substitute):
    """Addons>SelectValue (custom_ops/select_value/select_value_ops.cc:33-56): values outside the
    interval set are replaced by `substitute`."""
    return np.where(_in_intervals(values, intervals), values, np.asarray(substitute, values.dtype))


def gather_str_value_gen_indice(strings, not_equal: bytes, dense_rank: int = 2):
    """Addons>GatherStrValueGenIndice (custom_ops/gather_str_value_gen_indice/
    gather_str_value_gen_indice_ops.cc:31-58): (strings [n]) -> (indices [m, dense_rank] = [i, 0..],
    the strings != not_equal)."""
    keep = [i for i, s in enumerate(strings) if s != not_equal]
    idx = np.zeros((len(keep), dense_rank), np.int64)
    idx[:, 0] = keep
    return idx, [strings[i] for i in keep]


def string_split(strings, delimiter: bytes, skip_empty: bool = True):
    """tf.strings.split / StringSplit restated (TF/tensorflow/core/kernels/string_split_op.cc:37-99:
    SplitOnChar / SplitOnCharSet / Split; :162-170 skip_empty defaults to true). An empty string has
    no tokens; an empty delimiter splits character by character; otherwise ANY character of
    `delimiter` splits, and empty tokens are dropped when skip_empty.
    -> (tokens: list of bytes, row_splits int32[len(strings) + 1])."""
    tokens, rs = [], [0]
    for s in strings:
        s = bytes(s)
        if len(s) == 0:
            pass
        elif len(delimiter) == 0:
            tokens.extend(bytes([c]) for c in s)
        else:
            start = 0
            for i in range(len(s) + 1):
                if i == len(s) or s[i] in delimiter:
                    tok = s[start:i]
                    if not skip_empty or len(tok) > 0:
                        tokens.append(tok)
                    start = i + 1
        rs.append(len(tokens))
    return tokens, np.asarray(rs, np.int32)
