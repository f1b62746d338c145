"""Column descriptors and the request wire format (host side, numpy only).

A `Model` is what the reference's code generator knows about a graph after its
passes ran (tensorflow_addons/graph_optimizers/cuda_emitter.cc:975-1038): for every
embedding column its transform, lookup kind, table and concat slot. Here that
knowledge is data (`recom_column_desc`, include/recom_b200.h), not generated code.

A `PackedRequest` is the blob `Addons>ConcatInputs` produces
(tensorflow_addons/custom_ops/concat_inputs/concat_inputs_ops.cc:42-77): every
host-produced input of every column copied into one buffer (one H2D copy), plus
per-slot offsets and element counts. Differences from the reference's blob: slots
are 16-byte aligned, string tensors travel as (int32 offsets[n+1], bytes), and
multi-hot row membership can be sent as int32 row_splits[B+1] instead of the
int64 [nnz, rank] SparseTensor.indices.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import consts as L

NP_OF = {L.F32: np.float32, L.I32: np.int32, L.I64: np.int64, L.F64: np.float64}


@dataclass
class Filter:
    """Pre-lookup id filter of a column (include/recom_b200.h: recom_column_filter). kind KEEP /
    SELECT: `intervals` = closed [lo, hi] pairs over the raw integer values; SELECT replaces the
    values outside by `substitute`, KEEP drops them. STR_NE drops the strings equal to `not_equal`."""

    kind: int
    intervals: Sequence[Tuple[int, int]] = ()
    substitute: int = 0
    not_equal: bytes = b""


@dataclass
class Column:
    """One embedding column (the reference's "FC", graph_info.h:41-42)."""

    name: str
    lookup: int
    transform: int
    value_dtype: int
    table: int
    table_rows: int
    dim: int
    combiner: int = L.COMBINE_SUM
    boundaries: Optional[np.ndarray] = None  # float32, BUCKETIZE only
    num_buckets: int = 0  # HASH_* only
    seg_format: int = L.SEG_NONE
    seg_stride: int = 2  # COO: int64s per entry; SPLIT_STR: consts.split_spec(delimiter, skip_empty)
    input_key: Optional[str] = None  # columns with the same key share input slots
    filter: Optional[Filter] = None  # pre-lookup id filter
    # assigned by Model.finalize()
    out_offset: int = -1
    value_slot: int = -1
    bytes_slot: int = -1
    seg_slot: int = -1

    @property
    def key(self) -> str:
        return self.input_key or self.name


@dataclass
class Model:
    columns: List[Column]
    table_shapes: List[Tuple[int, int]]  # (rows, dim) per table
    out_dtype: int = L.F32
    name: str = "model"
    out_width: int = 0
    num_slots: int = 0
    slot_of_key: Dict[str, Tuple[int, int, int]] = field(default_factory=dict)

    def finalize(self) -> "Model":
        """Assign concat offsets (column order == ConcatV2 input order,
        concat_outputs_op_gpu.cu.cc:74-80) and blob slots."""
        off = 0
        slot = 0
        self.slot_of_key = {}
        for c in self.columns:
            c.out_offset = off
            off += c.dim
            if c.key in self.slot_of_key:
                c.value_slot, c.bytes_slot, c.seg_slot = self.slot_of_key[c.key]
                continue
            c.value_slot = slot
            slot += 1
            c.bytes_slot = -1
            if c.value_dtype == L.STR:
                c.bytes_slot = slot
                slot += 1
            c.seg_slot = -1
            if c.lookup != L.LOOKUP_DENSE and c.seg_format not in (L.SEG_SPLIT_STR, L.SEG_FIXED):
                c.seg_slot = slot
                slot += 1
                if c.seg_format == L.SEG_COO_RESHAPE_I64:  # dense_shape (+ new shape) behind the indices
                    slot += 2 if c.seg_stride & 256 else 1
            self.slot_of_key[c.key] = (c.value_slot, c.bytes_slot, c.seg_slot)
        self.out_width = off
        self.num_slots = slot
        return self

    def descs(self) -> Tuple[np.ndarray, np.ndarray]:
        """-> (recom_column_desc[num_cols], float32 boundary array)."""
        d = np.zeros(len(self.columns), dtype=L.COLUMN_DESC)
        bnd: List[np.ndarray] = []
        boff = 0
        table_dtype = self.out_dtype
        for i, c in enumerate(self.columns):
            d[i]["lookup"] = c.lookup
            d[i]["combiner"] = c.combiner
            d[i]["transform"] = c.transform
            d[i]["value_dtype"] = c.value_dtype
            d[i]["table"] = c.table
            d[i]["table_dtype"] = table_dtype
            d[i]["table_rows"] = c.table_rows
            d[i]["dim"] = c.dim
            d[i]["out_offset"] = c.out_offset
            d[i]["value_slot"] = c.value_slot
            d[i]["bytes_slot"] = c.bytes_slot
            d[i]["seg_slot"] = c.seg_slot
            d[i]["seg_format"] = c.seg_format if c.lookup != L.LOOKUP_DENSE else L.SEG_NONE
            d[i]["seg_stride"] = c.seg_stride
            d[i]["num_buckets"] = c.num_buckets
            if c.transform == L.XFORM_BUCKETIZE:
                b = np.ascontiguousarray(c.boundaries, dtype=np.float32)
                d[i]["num_boundaries"] = b.size
                d[i]["boundaries_offset"] = boff
                bnd.append(b)
                boff += b.size
        boundaries = np.concatenate(bnd) if bnd else np.zeros(0, np.float32)
        return d, boundaries

    def filters(self) -> Optional[np.ndarray]:
        """-> recom_column_filter[num_cols], or None when no column has a filter."""
        if not any(c.filter for c in self.columns):
            return None
        f = np.zeros(len(self.columns), dtype=L.COLUMN_FILTER)
        for i, c in enumerate(self.columns):
            if not c.filter:
                continue
            f[i]["kind"] = c.filter.kind
            iv = list(c.filter.intervals)
            f[i]["num_intervals"] = len(iv)
            for k, (lo, hi) in enumerate(iv):
                f[i]["lo"][k], f[i]["hi"][k] = lo, hi
            f[i]["substitute"] = c.filter.substitute
            ne = bytes(c.filter.not_equal)
            f[i]["not_equal_len"] = len(ne)
            f[i]["not_equal"][:len(ne)] = np.frombuffer(ne, np.uint8)
        return f

    def table_bytes(self) -> int:
        es = 2 if self.out_dtype == L.BF16 else 4
        return sum(r * d * es for r, d in self.table_shapes)


@dataclass
class Strings:
    """Packed string tensor: offsets int32[n+1] into a uint8 payload."""

    offsets: np.ndarray
    data: np.ndarray

    @staticmethod
    def from_list(items: Sequence[Union[bytes, str]]) -> "Strings":
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in items]
        off = np.zeros(len(bs) + 1, np.int32)
        if bs:
            off[1:] = np.cumsum([len(b) for b in bs])
        data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
        return Strings(off, data)

    def __len__(self) -> int:
        return self.offsets.size - 1

    def tolist(self) -> List[bytes]:
        raw = self.data.tobytes()
        return [raw[self.offsets[i]:self.offsets[i + 1]] for i in range(len(self))]


@dataclass
class Feature:
    """Host-side input of one column (or of a group sharing `input_key`).

    values: ndarray (f32/i32/i64/f64) or Strings. For DENSE columns len == batch.
    row_splits: int32[batch+1] (SEG_ROW_SPLITS_I32) or
    coo: int64[nnz, stride] SparseTensor.indices (SEG_COO_I64).
    """

    values: Union[np.ndarray, Strings]
    row_splits: Optional[np.ndarray] = None
    coo: Optional[np.ndarray] = None
    dense_shape: Optional[np.ndarray] = None  # SEG_COO_RESHAPE_I64: int64[rank] of the SparseTensor
    new_shape: Optional[np.ndarray] = None    # ... and the SparseReshape target, when the column gives one


@dataclass
class PackedRequest:
    batch: int
    blob: np.ndarray  # uint8
    slot_offsets: np.ndarray  # int64[num_slots]
    slot_elems: np.ndarray  # int64[num_slots]
    total_ids: int  # ids over all columns (ids_out length)
    sparse_ids: int  # ids over non-DENSE columns (workspace sizing)
    ids_base: np.ndarray  # int64[num_cols+1]
    args_offset: int = -1  # byte offset of the argument trailer inside the blob, -1: none


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


ARGS_MAGIC = 0x5247524132304252  # kArgsMagic (csrc/plan.h)


def pack_request(model: Model, batch: int, features: Dict[str, Feature],
                 pinned_blob: Optional[np.ndarray] = None, trailer: bool = True) -> PackedRequest:
    """numpy mirror of Addons>ConcatInputs (the C++ one: recom_concat_inputs_with_args in
    csrc/host_ops.cc). trailer: append {magic, num_slots, slot_offsets[], slot_elems[]} (int64)
    behind the last slot, so that the blob's one H2D copy carries the request's argument data
    and recom_plan_run uploads nothing."""
    slot_arrays: List[Optional[np.ndarray]] = [None] * model.num_slots
    done = set()
    for c in model.columns:
        if c.key in done:
            continue
        done.add(c.key)
        f = features[c.key]
        if c.value_dtype == L.STR:
            assert isinstance(f.values, Strings), c.name
            slot_arrays[c.value_slot] = np.ascontiguousarray(f.values.offsets, np.int32)
            slot_arrays[c.bytes_slot] = np.ascontiguousarray(f.values.data, np.uint8)
        else:
            slot_arrays[c.value_slot] = np.ascontiguousarray(f.values, NP_OF[c.value_dtype])
        if c.lookup != L.LOOKUP_DENSE and c.seg_format not in (L.SEG_SPLIT_STR, L.SEG_FIXED):
            if c.seg_format == L.SEG_ROW_SPLITS_I32:
                assert f.row_splits is not None, c.name
                slot_arrays[c.seg_slot] = np.ascontiguousarray(f.row_splits, np.int32)
            elif c.seg_format == L.SEG_COO_RESHAPE_I64:
                assert f.coo is not None and f.dense_shape is not None, c.name
                slot_arrays[c.seg_slot] = np.ascontiguousarray(f.coo, np.int64).reshape(-1)
                slot_arrays[c.seg_slot + 1] = np.ascontiguousarray(f.dense_shape, np.int64)
                if c.seg_stride & 256:
                    assert f.new_shape is not None, c.name
                    slot_arrays[c.seg_slot + 2] = np.ascontiguousarray(f.new_shape, np.int64)
            else:
                assert f.coo is not None, c.name
                slot_arrays[c.seg_slot] = np.ascontiguousarray(f.coo, np.int64).reshape(-1)
    offsets = np.zeros(model.num_slots, np.int64)
    elems = np.zeros(model.num_slots, np.int64)
    off = 0
    for s, a in enumerate(slot_arrays):
        assert a is not None, f"slot {s} has no producer"
        offsets[s] = off
        elems[s] = a.size
        off = _align(off + a.nbytes)
    data_end = _align(max(off, 16))
    total = data_end + (_align(16 + 16 * model.num_slots) if trailer else 0)
    if pinned_blob is not None:
        assert pinned_blob.nbytes >= total
        blob = pinned_blob[:total]
    else:
        blob = np.zeros(total, np.uint8)
    for s, a in enumerate(slot_arrays):
        blob[offsets[s]:offsets[s] + a.nbytes] = a.view(np.uint8).reshape(-1)
    if trailer:
        t = np.concatenate([np.array([ARGS_MAGIC, model.num_slots], np.int64), offsets, elems])
        blob[data_end:data_end + t.nbytes] = t.view(np.uint8)
    ids_base = np.zeros(len(model.columns) + 1, np.int64)
    sparse = 0
    for i, c in enumerate(model.columns):
        if c.lookup == L.LOOKUP_DENSE:
            n = batch
        elif c.seg_format == L.SEG_SPLIT_STR:
            # tokens are counted on the device; the workspace holds payload bytes + batch + 1 ids
            n = 0
            sparse += int(elems[c.bytes_slot]) + batch + 1
        else:
            n = int(elems[c.value_slot]) - (1 if c.value_dtype == L.STR else 0)
            sparse += n
        ids_base[i + 1] = ids_base[i] + n
    return PackedRequest(batch, blob, offsets, elems, int(ids_base[-1]), sparse, ids_base,
                         data_end if trailer else -1)


def algorithmic_bytes(model: Model, req: PackedRequest) -> int:
    """sum_c [in_c + seg_c + tbl_c + out_c]; python mirror of
    recom_plan_algorithmic_bytes (used to cross-check it in tests)."""
    es = 2 if model.out_dtype == L.BF16 else 4
    total = 0
    for i, c in enumerate(model.columns):
        nnz = int(req.ids_base[i + 1] - req.ids_base[i])
        if c.lookup != L.LOOKUP_DENSE and c.seg_format == L.SEG_SPLIT_STR:
            # mirror of the C side: offsets + payload, no table term (token count unknown on the host)
            total += 4 * (req.batch + 1) + int(req.slot_elems[c.bytes_slot]) + c.dim * es * req.batch
            continue
        if c.value_dtype == L.STR:
            inb = 4 * (nnz + 1) + int(req.slot_elems[c.bytes_slot])
        else:
            inb = L.DTYPE_BYTES[c.value_dtype] * nnz
        seg = 0 if c.lookup == L.LOOKUP_DENSE else 4 * (req.batch + 1)
        tbl = c.dim * es * min(nnz, c.table_rows)
        total += inb + seg + tbl + c.dim * es * req.batch
    return total
