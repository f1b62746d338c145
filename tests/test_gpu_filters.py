"""SURVEY 8(f) N2 - pre-lookup id filtering fused into the stage. The GPU path runs the model WITH
the filters on the raw features; the expected result is the oracle on the model WITHOUT filters,
fed the features the reference's three ops would have produced (numpy restatements in
oracle/oracle.py of gather_indice_value_ops.cc:33-70, select_value_ops.cc:33-56,
gather_str_value_gen_indice_ops.cc:31-58)."""
import copy

import numpy as np
import pytest

from oracle import oracle as O
from recom_b200 import _lib as L
from recom_b200 import synth
from recom_b200.schema import Column, Feature, Filter, Model, Strings, pack_request
from tests.helpers import run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def _model(filters: bool):
    F = (lambda f: f) if filters else (lambda f: None)
    cols = [
        # multi-hot raw ids, drop the negative ones (sparse.retain(ids, ids >= 0)): mean over the kept ids
        Column("keep_mh", L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I64, 0, 64, 8, combiner=L.COMBINE_MEAN,
               seg_format=L.SEG_ROW_SPLITS_I32, filter=F(Filter(L.FILTER_KEEP, [(0, 63)]))),
        # the same input once more through two intervals, sqrtn
        Column("keep2_mh", L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I64, 1, 64, 16, combiner=L.COMBINE_SQRTN,
               seg_format=L.SEG_ROW_SPLITS_I32, filter=F(Filter(L.FILTER_KEEP, [(0, 9), (20, 40)]))),
        # one id per row, out-of-range values replaced by the default id 3
        Column("select_dense", L.LOOKUP_DENSE, L.XFORM_NONE, L.I32, 2, 50, 8,
               filter=F(Filter(L.FILTER_SELECT, [(0, 49)], substitute=3))),
        # hashed integers, values outside [10, 90] mapped to 0 first
        Column("select_hash", L.LOOKUP_SEGMENT, L.XFORM_HASH_INT, L.I64, 3, 97, 4, combiner=L.COMBINE_SUM,
               num_buckets=97, seg_format=L.SEG_ROW_SPLITS_I32,
               filter=F(Filter(L.FILTER_SELECT, [(10, 90)], substitute=0))),
        # strings, "" dropped (the default_value of a dense string feature)
        Column("str_ne", L.LOOKUP_SEGMENT, L.XFORM_HASH_STR, L.STR, 4, 101, 8, combiner=L.COMBINE_MEAN,
               num_buckets=101, seg_format=L.SEG_ROW_SPLITS_I32,
               filter=F(Filter(L.FILTER_STR_NE, not_equal=b""))),
        # one id per row, dropped -> zero row
        Column("keep_dense", L.LOOKUP_DENSE, L.XFORM_NONE, L.I64, 5, 30, 4,
               filter=F(Filter(L.FILTER_KEEP, [(0, 29)]))),
    ]
    tables = [(64, 8), (64, 16), (50, 8), (97, 4), (101, 8), (30, 4)]
    return Model(cols, tables, name="filtered" if filters else "plain").finalize()


def _features(batch, seed):
    rng = np.random.default_rng(seed)

    def mh(lo, hi, maxlen=12, long_row=None):
        n = rng.integers(0, maxlen, size=batch)
        if long_row is not None:
            n[long_row] = 150  # a row handled by all lane groups together
        rs = np.zeros(batch + 1, np.int32)
        rs[1:] = np.cumsum(n)
        return rng.integers(lo, hi, size=int(rs[-1])), rs

    f = {}
    v, rs = mh(-20, 64, long_row=batch // 2)
    f["keep_mh"] = Feature(v.astype(np.int64), row_splits=rs)
    v, rs = mh(-5, 64)
    f["keep2_mh"] = Feature(v.astype(np.int64), row_splits=rs)
    f["select_dense"] = Feature(rng.integers(-10, 70, size=batch).astype(np.int32))
    v, rs = mh(0, 120)
    f["select_hash"] = Feature(v.astype(np.int64), row_splits=rs)
    n = rng.integers(0, 6, size=batch)
    rs = np.zeros(batch + 1, np.int32)
    rs[1:] = np.cumsum(n)
    strs = [b"" if rng.random() < 0.3 else bytes(rng.choice(list(b"abcdefgh"), size=rng.integers(1, 14)).tolist())
            for _ in range(int(rs[-1]))]
    f["str_ne"] = Feature(Strings.from_list(strs), row_splits=rs)
    f["keep_dense"] = Feature(rng.integers(-5, 40, size=batch).astype(np.int64))
    return f


def _apply_reference_ops(model, feats, batch):
    """What the reference's TF-CPU filter ops hand to the lookup: the filtered features."""
    out = {}
    for c in model.columns:
        ft, flt = feats[c.name], c.filter
        if flt.kind == L.FILTER_SELECT:
            out[c.name] = Feature(O.select_value(ft.values, flt.intervals, flt.substitute), row_splits=ft.row_splits)
            continue
        if c.lookup == L.LOOKUP_DENSE:
            # one id per row: as a sparse column [row, 0] -> dropped rows become empty rows
            idx = np.stack([np.arange(batch, dtype=np.int64), np.zeros(batch, np.int64)], 1)
            idx2, vals2 = O.gather_indice_value(idx, ft.values, flt.intervals)
            rs = np.zeros(batch + 1, np.int32)
            rs[1:] = np.cumsum(np.bincount(idx2[:, 0], minlength=batch))
            out[c.name] = Feature(vals2, row_splits=rs)
            continue
        rows = np.repeat(np.arange(batch, dtype=np.int64), np.diff(ft.row_splits))
        if flt.kind == L.FILTER_STR_NE:
            strs = ft.values.tolist()
            idx2, kept = O.gather_str_value_gen_indice(strs, flt.not_equal)
            rows2 = rows[idx2[:, 0]]
            vals2 = Strings.from_list(kept)
        else:
            idx = np.stack([rows, np.zeros_like(rows)], 1)
            idx2, vals2 = O.gather_indice_value(idx, ft.values, flt.intervals)
            rows2 = idx2[:, 0]
        rs = np.zeros(batch + 1, np.int32)
        rs[1:] = np.cumsum(np.bincount(rows2, minlength=batch))
        out[c.name] = Feature(vals2, row_splits=rs)
    return out


@pytest.mark.parametrize("batch,seed", [(257, 0), (40, 3)])
def test_fused_filters_match_reference_ops(batch, seed):
    with_f, plain = _model(True), _model(False)
    feats = _features(batch, seed)
    tables = synth.make_tables_np(with_f, seed=seed)
    got, _, status, _ = run_gpu(with_f, pack_request(with_f, batch, feats), tables, want_ids=False)
    assert status == 0
    # expected: the plain model on the features the reference's ops produce; a one-id-per-row
    # column whose id was dropped is an empty row of a multi-hot (sum) column: zeros
    exp_model = copy.deepcopy(plain)
    for c, cf in zip(exp_model.columns, with_f.columns):
        if c.lookup == L.LOOKUP_DENSE and cf.filter.kind == L.FILTER_KEEP:
            c.lookup, c.combiner, c.seg_format = L.LOOKUP_SEGMENT, L.COMBINE_SUM, L.SEG_ROW_SPLITS_I32
    exp_model = Model(exp_model.columns, exp_model.table_shapes, name="expected").finalize()
    ref, _ = run_oracle(exp_model, pack_request(exp_model, batch, _apply_reference_ops(with_f, feats, batch)), tables)
    err = np.abs(got - ref)
    assert np.all(err <= 1e-5 * np.maximum(np.abs(ref), 1.0)), float(err.max())
    # sanity: the filters do change the result
    got_plain, _, _, _ = run_gpu(plain, pack_request(plain, batch, _features(batch, seed)), tables, want_ids=False,
                                 check_ids=True)
    assert np.abs(got_plain - got).max() > 1e-3


def test_filtered_plan_survives_serialization():
    from recom_b200 import runtime as R

    m = _model(True)
    plan = R.Plan(m)
    blob = plan.serialize()
    plan2 = R.Plan.from_blob(blob, m)
    assert plan2.serialize() == blob
    plan.close()
    plan2.close()
