"""Synthetic models and feature generators of the reference's benchmark configs.

Shapes follow the reference's artifact-evaluation scripts (reference root relative):
  * model E / F       examples/python/dlrm.py:31-37 (sizes), :56-137 (column makers),
                      :202-203 (column counts); concat order = sorted column names, as
                      tf.feature_column.input_layer sorts its columns by name.
  * microbenchmark    examples/python/microbenchmark.py:31-66.
  * C++ harness inputs examples/cc/recom_examples.patch:3332-3355 (decimal strings,
                      1..10 ids per sample).
Plus the BASELINE.json extensions: mixed dims / combiners, dynamic shape, large-table,
production-shaped (shared tables and duplicated inputs).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import consts as L
from .schema import Column, Feature, Model, Strings

ALPHABET = np.frombuffer(
    b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789", dtype=np.uint8)

SMALL_ROWS = 100          # dlrm.py:31
MID_ROWS = 10000          # dlrm.py:32
LARGE_ROWS = 1 << 23      # dlrm.py:33
EMB_DIM = 8               # dlrm.py:34
LARGE_DIM = 32            # dlrm.py:35
MAX_IDS = 10              # dlrm.py:37 input_cols


def _bucketize_col(name, table, rows=SMALL_ROWS, dim=EMB_DIM, step=5, dtype=L.F32):
    # dlrm.py:124-137 / microbenchmark.py:46-58: boundaries 0,step,..; table rows+1
    b = np.arange(0, rows * step, step, dtype=np.float32)
    return Column(name, L.LOOKUP_DENSE, L.XFORM_BUCKETIZE, dtype, table, rows + 1, dim,
                  combiner=L.COMBINE_MEAN, boundaries=b)


def _hash_int_col(name, table, rows=SMALL_ROWS, dim=EMB_DIM):
    # dlrm.py:56-70
    return Column(name, L.LOOKUP_DENSE, L.XFORM_HASH_INT, L.I32, table, rows, dim,
                  combiner=L.COMBINE_MEAN, num_buckets=rows)


def _hash_str_col(name, table, rows=MID_ROWS, dim=EMB_DIM):
    # dlrm.py:73-87
    return Column(name, L.LOOKUP_DENSE, L.XFORM_HASH_STR, L.STR, table, rows, dim,
                  combiner=L.COMBINE_MEAN, num_buckets=rows)


def _multihot_col(name, table, rows, dim, combiner=L.COMBINE_SUM, seg_format=L.SEG_ROW_SPLITS_I32,
                  hashed=True):
    # dlrm.py:90-121
    if hashed:
        return Column(name, L.LOOKUP_SEGMENT, L.XFORM_HASH_STR, L.STR, table, rows, dim,
                      combiner=combiner, num_buckets=rows, seg_format=seg_format)
    return Column(name, L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I64, table, rows, dim,
                  combiner=combiner, seg_format=seg_format)


def dlrm_model(n_bucketize: int, n_int: int, n_str: int, n_sparse: int, n_large: int,
               large_rows: int = LARGE_ROWS, out_dtype: int = L.F32, name: str = "dlrm",
               seg_format: int = L.SEG_ROW_SPLITS_I32, fused_hash: bool = True) -> Model:
    """create_dlrm (dlrm.py:140-166). fused_hash=False feeds pre-hashed int64 ids, which is
    what the reference's op receives (hashing stays on TF-CPU there)."""
    named: List[Tuple[str, str]] = []
    named += [(f"bucketize{i}", "b") for i in range(n_bucketize)]
    named += [(f"categ_hashbucket_int{i}", "i") for i in range(n_int)]
    named += [(f"categ_hashbucket{i}", "s") for i in range(n_str)]
    named += [(f"categ_hashbucket_sparse{i}", "m") for i in range(n_sparse)]
    named += [(f"categ_large_hashbucket_sparse{i}", "l") for i in range(n_large)]
    named.sort(key=lambda t: t[0])  # input_layer sorts columns by name
    cols: List[Column] = []
    tables: List[Tuple[int, int]] = []
    for cname, kind in named:
        t = len(tables)
        if kind == "b":
            c = _bucketize_col(cname, t)
        elif kind == "i":
            c = _hash_int_col(cname, t)
        elif kind == "s":
            c = _hash_str_col(cname, t)
        elif kind == "m":
            c = _multihot_col(cname, t, MID_ROWS, EMB_DIM, seg_format=seg_format, hashed=fused_hash)
        else:
            c = _multihot_col(cname, t, large_rows, LARGE_DIM, seg_format=seg_format,
                              hashed=fused_hash)
        if not fused_hash and kind in ("i", "s"):
            c.transform, c.value_dtype, c.num_buckets = L.XFORM_NONE, L.I64, 0
        cols.append(c)
        tables.append((c.table_rows, c.dim))
    return Model(cols, tables, out_dtype=out_dtype, name=name).finalize()


def model_E(**kw) -> Model:
    return dlrm_model(880, 50, 50, 15, 5, name="E", **kw)  # dlrm.py:202


def model_F(**kw) -> Model:
    return dlrm_model(1000, 90, 100, 7, 3, name="F", **kw)  # dlrm.py:203


def model_E_small(**kw) -> Model:
    """Config 1(ii): E/10."""
    kw.setdefault("large_rows", 1 << 16)
    return dlrm_model(88, 5, 5, 2, 0, name="E_div10", **kw)


def microbenchmark_model(n: int = 100, rows: int = 100, dim: int = 8,
                         random_boundary: bool = False, seed: int = 0) -> Model:
    """microbenchmark.py:31-66: N bucketized + embedding(mean) columns."""
    rng = np.random.default_rng(seed)
    cols, tables = [], []
    for i in range(n):
        r, step = rows, 5
        if random_boundary:  # microbenchmark.py:54-57
            step = int(rng.integers(5, 11))
            r = rows + int(rng.integers(-50, 51))
        cols.append(_bucketize_col(f"f{i}", i, rows=r, dim=dim, step=step))
        tables.append((r + 1, dim))
    cols.sort(key=lambda c: c.name)
    return Model(cols, tables, name=f"micro{n}").finalize()


def mixed_model(n_cols: int = 1000, seed: int = 0, out_dtype: int = L.F32,
                large_rows: int = 1 << 20, frac_multihot: float = 0.1) -> Model:
    """BASELINE.json config 2 extension: dims from {8,16,32,64}, combiners from
    {sum, mean, sqrtn}, a mix of transforms and lookup kinds."""
    rng = np.random.default_rng(seed)
    cols, tables = [], []
    for i in range(n_cols):
        dim = int(rng.choice([8, 16, 32, 64], p=[0.6, 0.2, 0.15, 0.05]))
        u = rng.random()
        t = len(tables)
        if u < frac_multihot:
            rows = int(rng.choice([MID_ROWS, large_rows]))
            comb = int(rng.choice([L.COMBINE_SUM, L.COMBINE_MEAN, L.COMBINE_SQRTN]))
            fmt = int(rng.choice([L.SEG_ROW_SPLITS_I32, L.SEG_COO_I64]))
            c = _multihot_col(f"mh{i}", t, rows, dim, combiner=comb, seg_format=fmt,
                              hashed=bool(rng.random() < 0.7))
        elif u < frac_multihot + 0.05:
            c = Column(f"gs{i}", L.LOOKUP_GATHER_SCATTER, L.XFORM_NONE, L.I64, t, MID_ROWS, dim,
                       seg_format=L.SEG_COO_I64, seg_stride=int(rng.choice([1, 2])))
        elif u < 0.8:
            c = _bucketize_col(f"bk{i}", t, rows=int(rng.integers(20, 200)), dim=dim,
                               step=int(rng.integers(1, 10)),
                               dtype=int(rng.choice([L.F32, L.I32, L.I64])))
        elif u < 0.9:
            c = _hash_int_col(f"hi{i}", t, rows=int(rng.integers(50, 5000)), dim=dim)
        else:
            c = _hash_str_col(f"hs{i}", t, rows=int(rng.integers(100, 50000)), dim=dim)
        cols.append(c)
        tables.append((c.table_rows, c.dim))
    return Model(cols, tables, out_dtype=out_dtype, name=f"mixed{n_cols}").finalize()


def large_table_model(n_cols: int = 1000, rows: int = 1_500_000, dim: int = 32,
                      out_dtype: int = L.F32) -> Model:
    """BASELINE.json config 4: n_cols multi-hot mean columns over [rows, dim] tables
    (1000 x [1.5M, 32] fp32 = 192 GB aggregate); ids arrive pre-hashed."""
    cols = [_multihot_col(f"lt{i:04d}", i, rows, dim, combiner=L.COMBINE_MEAN, hashed=False)
            for i in range(n_cols)]
    return Model(cols, [(rows, dim)] * n_cols, out_dtype=out_dtype, name="large_table").finalize()


def production_model(n_cols: int = 2000, share: int = 4, seed: int = 0) -> Model:
    """BASELINE.json config 5: groups of `share` columns alias one table, and every
    other group also re-reads the same input (duplicate (input, transform) pairs)."""
    rng = np.random.default_rng(seed)
    cols, tables = [], []
    i = 0
    while i < n_cols:
        t = len(tables)
        kind = rng.random()
        dup_input = bool(rng.random() < 0.5)
        group = min(share, n_cols - i)
        if kind < 0.8:
            proto = _bucketize_col("x", t)
        elif kind < 0.9:
            proto = _hash_str_col("x", t)
        else:
            proto = _multihot_col("x", t, MID_ROWS, EMB_DIM, combiner=L.COMBINE_MEAN)
        tables.append((proto.table_rows, proto.dim))
        for g in range(group):
            c = Column(**{**proto.__dict__})
            c.name = f"p{i:05d}"
            c.input_key = f"in{t}" if dup_input else None
            cols.append(c)
            i += 1
    return Model(cols, tables, name=f"prod{n_cols}").finalize()


# --------------------------------------------------------------------------------
# features
# --------------------------------------------------------------------------------
def random_strings(rng: np.random.Generator, n: int, length: int = 8) -> Strings:
    """dlrm.py:44-46: n strings of `length` chars from [A-Za-z0-9]."""
    data = ALPHABET[rng.integers(0, ALPHABET.size, size=n * length)]
    off = (np.arange(n + 1, dtype=np.int64) * length).astype(np.int32)
    return Strings(off, np.ascontiguousarray(data))


def decimal_strings(rng: np.random.Generator, n: int) -> Strings:
    """recom_examples.patch:3340-3343: to_string(rand 0..100) cut 4 digits after '.'."""
    vals = rng.random(n) * 100.0
    return Strings.from_list([f"{v:.6f}"[: f"{v:.6f}".find(".") + 5] for v in vals])


def row_lengths(rng: np.random.Generator, batch: int, dist: str = "u10",
                empty_frac: float = 0.0) -> np.ndarray:
    if dist == "u10":      # dlrm.py:52 random.randint(1, 10)
        n = rng.integers(1, MAX_IDS + 1, size=batch)
    elif dist == "u50":
        n = rng.integers(1, 51, size=batch)
    elif dist == "zipf":
        n = np.minimum(rng.zipf(1.5, size=batch), 2000)
    elif dist == "one":
        n = np.ones(batch, np.int64)
    else:
        raise ValueError(dist)
    if empty_frac > 0:
        n = np.where(rng.random(batch) < empty_frac, 0, n)
    return n.astype(np.int64)


def splits_and_coo(lengths: np.ndarray, stride: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    rs = np.zeros(lengths.size + 1, np.int32)
    rs[1:] = np.cumsum(lengths)
    rows = np.repeat(np.arange(lengths.size, dtype=np.int64), lengths)
    coo = np.zeros((rows.size, stride), np.int64)
    coo[:, 0] = rows
    if stride > 1:
        coo[:, 1] = np.arange(rows.size) - np.repeat(rs[:-1].astype(np.int64), lengths)
    return rs, coo


def make_features(model: Model, batch: int, seed: int = 0, dist: str = "u10",
                  empty_frac: float = 0.0, shared_lengths: bool = False) -> Dict[str, Feature]:
    """Random inputs with the distributions of dlrm.py:40-53. `shared_lengths` mimics the
    C++ harness, where all multi-hot columns of a sample share one length
    (recom_examples.patch:3369-3370)."""
    rng = np.random.default_rng(seed)
    feats: Dict[str, Feature] = {}
    common = row_lengths(rng, batch, dist, empty_frac) if shared_lengths else None
    for c in model.columns:
        if c.key in feats:
            continue
        if c.lookup == L.LOOKUP_DENSE:
            n, rs, coo = batch, None, None
        else:
            if c.lookup == L.LOOKUP_GATHER_SCATTER:
                lengths = (rng.random(batch) < 0.7).astype(np.int64)
            else:
                lengths = common if common is not None else row_lengths(rng, batch, dist, empty_frac)
            rs, coo = splits_and_coo(lengths, c.seg_stride)
            n = int(rs[-1])
        if c.transform == L.XFORM_BUCKETIZE:
            hi = float(c.boundaries[-1]) + 10 if c.boundaries.size else 10.0
            if c.value_dtype in (L.F32, L.F64):
                # dlrm.py:40-41 feeds integers cast to float; add fractional and
                # out-of-range values so every bucket edge is exercised
                v = rng.integers(-5, int(hi), size=n).astype(np.float64)
                v += np.where(rng.random(n) < 0.5, 0.0, rng.random(n))
                vals = v.astype(np.float32 if c.value_dtype == L.F32 else np.float64)
            else:
                vals = rng.integers(-5, int(hi), size=n).astype(
                    np.int32 if c.value_dtype == L.I32 else np.int64)
        elif c.transform == L.XFORM_HASH_INT:
            vals = rng.integers(0, 100, size=n).astype(np.int32 if c.value_dtype == L.I32 else np.int64)
        elif c.transform == L.XFORM_HASH_STR:
            vals = random_strings(rng, n, 8)
        else:
            vals = rng.integers(0, c.table_rows, size=n).astype(
                np.int32 if c.value_dtype == L.I32 else np.int64)
        feats[c.key] = Feature(vals, row_splits=rs, coo=coo)
    return feats


def make_tables_np(model: Model, seed: int = 0, as_bf16_bits: bool = False) -> List[np.ndarray]:
    """Host tables for small models: N(0, 1/sqrt(D)) like TF's embedding initializer."""
    rng = np.random.default_rng(seed + 12345)
    out = []
    for rows, dim in model.table_shapes:
        t = (rng.standard_normal((rows, dim)) / np.sqrt(dim)).astype(np.float32)
        out.append(t)
    return out
