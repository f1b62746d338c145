"""column descriptors extracted from a serialized GraphDef (recom_b200/graphdef.py)
against hand-assembled graphs of the input_layer shape (tests/graph_fixtures.py), and the
extracted model run through the oracle / the GPU like any hand-written one."""
import numpy as np
import pytest

from recom_b200 import consts as L
from recom_b200 import graphdef as G
from recom_b200 import synth
from recom_b200.schema import Feature, pack_request
from tests.graph_fixtures import Builder
from tests.helpers import run_oracle


def _dlrm_like():
    b = Builder()
    outs = [
        b.bucketized("bucketize0", np.arange(0, 500, 5), 8),
        b.hashed_dense("categ_hashbucket0", 10000, 8, string=True),
        b.hashed_dense("categ_hashbucket_int0", 100, 8, string=False),
        b.hashed_sparse("categ_hashbucket_sparse0", 10000, 8, rank=2, combiner="sum"),
        b.hashed_sparse("categ_large_hashbucket_sparse0", 1 << 14, 32, rank=2, combiner="sum"),
        b.identity_sparse("seq_ids", 777, 16, combiner="sqrtn"),
        b.hashed_sparse("rank3", 500, 8, rank=3, combiner="mean"),
    ]
    return b.finish(outs)


def test_wire_round_trip():
    data = _dlrm_like()
    g = G.parse_graphdef(data)
    assert G.encode_graphdef(g.nodes) == data  # reader and writer agree on every field they know
    bk = g.by_name["il/bucketize0/Bucketize"]
    assert bk.op == "Bucketize" and len(bk.attr["boundaries"]) == 100 and bk.attr["boundaries"][1] == 5.0
    ss = g.by_name["il/bucketize0_embedding/strided_slice"]
    assert ss.attr["shrink_axis_mask"] == 2 and ss.attr["begin_mask"] == 1
    np.testing.assert_array_equal(g.const_value(g.input_node(ss, 2)), [0, 1])
    assert g.by_name["categ_hashbucket0"].dtype() == G.DT_STRING


def test_tables_are_the_gather_only_variables():
    g = G.parse_graphdef(_dlrm_like())
    names = sorted(t.name for t in G.find_tables(g))
    assert len(names) == 7 and all(n.endswith("embedding_weights") for n in names)
    assert "dense/kernel" not in names  # consumed by MatMul: graph_info.cc:233-243


def test_extracted_descriptors():
    ex = G.extract_from_bytes(_dlrm_like())
    assert ex.skipped == []
    m = ex.model
    by = {c.name: c for c in m.columns}
    # slice order = ConcatV2 input order (cuda_emitter.cc:2556-2566)
    assert [c.name for c in m.columns] == ["bucketize0", "categ_hashbucket0", "categ_hashbucket_int0",
                                           "categ_hashbucket_sparse0/values",
                                           "categ_large_hashbucket_sparse0/values", "seq_ids/values", "rank3/values"]
    assert m.out_width == 8 + 8 + 8 + 8 + 32 + 16 + 8
    c = by["bucketize0"]
    assert (c.lookup, c.transform, c.value_dtype, c.table_rows, c.dim) == (L.LOOKUP_DENSE, L.XFORM_BUCKETIZE, L.F32, 101, 8)
    np.testing.assert_array_equal(c.boundaries, np.arange(0, 500, 5, dtype=np.float32))
    c = by["categ_hashbucket0"]
    assert (c.lookup, c.transform, c.value_dtype, c.num_buckets) == (L.LOOKUP_DENSE, L.XFORM_HASH_STR, L.STR, 10000)
    assert c.filter.kind == L.FILTER_STR_NE and c.filter.not_equal == b""
    c = by["categ_hashbucket_int0"]
    assert (c.transform, c.value_dtype, c.num_buckets) == (L.XFORM_HASH_INT, L.I32, 100)
    assert c.filter.kind == L.FILTER_KEEP
    c = by["categ_hashbucket_sparse0/values"]
    assert (c.lookup, c.seg_format, c.seg_stride, c.combiner) == (L.LOOKUP_SEGMENT, L.SEG_COO_I64, 2, L.COMBINE_SUM)
    c = by["seq_ids/values"]
    assert (c.transform, c.value_dtype, c.combiner, c.table_rows, c.dim) == (L.XFORM_NONE, L.I64, L.COMBINE_SQRTN, 777, 16)
    c = by["rank3/values"]
    assert (c.seg_format, c.seg_stride, c.combiner) == (L.SEG_COO_RESHAPE_I64, L.reshape_spec(3), L.COMBINE_MEAN)
    b = {x.column: x for x in ex.bindings}
    assert b["rank3/values"].indices == "rank3/indices" and b["rank3/values"].dense_shape == "rank3/shape"
    assert b["bucketize0"].table_variable == "il/bucketize0_embedding/embedding_weights"
    assert all(x.concat == "il/concat" for x in ex.bindings)


def test_unmatched_tables_are_reported():
    b = Builder()
    outs = [b.bucketized("f0", [0, 1, 2], 8)]
    # a table gathered directly (no safe_embedding_lookup_sparse chain around it)
    t = b.add("raw/embedding_weights", "VariableV2", shape=("shape", [10, 4]), dtype=("type", G.DT_FLOAT))
    ids = b.add("raw_ids", "Placeholder", dtype=("type", G.DT_INT64), shape=("shape", [-1]))
    outs.append(b.add("raw/GatherV2", "GatherV2", [t, ids, b.const("raw/axis", np.int32(0))]))
    ex = G.extract_from_bytes(b.finish(outs, with_mlp=False))
    assert [c.name for c in ex.model.columns] == ["f0"]
    assert ex.skipped and ex.skipped[0][0] == "raw/embedding_weights"


def _features_for(ex, batch, seed=0):
    rng = np.random.default_rng(seed)
    feats = {}
    for c in ex.model.columns:
        if c.lookup == L.LOOKUP_DENSE:
            if c.transform == L.XFORM_BUCKETIZE:
                feats[c.key] = Feature(rng.integers(-5, 520, size=batch).astype(np.float32))
            elif c.value_dtype == L.STR:
                s = synth.random_strings(rng, batch, 8).tolist()
                s = [b"" if rng.random() < 0.2 else x for x in s]  # the ignore value: dropped -> zero row
                feats[c.key] = Feature(synth.Strings.from_list(s))
            else:
                v = rng.integers(0, 100, size=batch).astype(np.int32)
                v[rng.random(batch) < 0.2] = -1
                feats[c.key] = Feature(v)
        elif c.seg_format == L.SEG_COO_I64:
            lengths = synth.row_lengths(rng, batch, "u10", 0.2)
            rs, coo = synth.splits_and_coo(lengths, 2)
            n = int(rs[-1])
            vals = synth.random_strings(rng, n, 8) if c.value_dtype == L.STR else rng.integers(0, c.table_rows, size=n).astype(np.int64)
            feats[c.key] = Feature(vals, coo=coo)
        else:  # rank 3 [batch / 4, 4, 6] -> [batch, 6]
            shape = (batch // 4, 4, 6)
            keep = np.flatnonzero(rng.random(int(np.prod(shape))) < 0.3)
            idx = np.stack(np.unravel_index(keep, shape), axis=1).astype(np.int64)
            feats[c.key] = Feature(synth.random_strings(rng, idx.shape[0], 8), coo=idx,
                                   dense_shape=np.array(shape, np.int64))
    return feats


def _expected(ex, batch, feats, tables):
    """The oracle on the extracted model without its filters, with the rows whose id the graph's
    Where(NotEqual(x, ignore)) drops zeroed (an empty row of safe_embedding_lookup_sparse)."""
    import copy

    plain = copy.deepcopy(ex.model)
    for c in plain.columns:
        c.filter = None
    req = pack_request(plain, batch, feats)
    out, _ = run_oracle(plain, req, tables)
    dropped = {}
    for c in ex.model.columns:
        if c.filter is None:
            continue
        v = feats[c.key].values
        drop = np.array([len(x) == 0 for x in v.tolist()]) if c.value_dtype == L.STR else (v == -1)
        out[drop, c.out_offset:c.out_offset + c.dim] = 0
        dropped[c.name] = drop
    return out, dropped


def test_extracted_model_runs_through_the_oracle():
    ex = G.extract_from_bytes(_dlrm_like())
    m, batch = ex.model, 64
    feats = _features_for(ex, batch)
    out, dropped = _expected(ex, batch, feats, synth.make_tables_np(m))
    assert out.shape == (batch, m.out_width) and np.isfinite(out).all()
    assert set(dropped) == {"categ_hashbucket0", "categ_hashbucket_int0"} and all(d.any() for d in dropped.values())
    for c in m.columns:  # every slice is written by its column
        assert (out[:, c.out_offset:c.out_offset + c.dim] != 0).any()


@pytest.mark.gpu
def test_extracted_model_gpu_parity():
    from tests.helpers import run_gpu

    ex = G.extract_from_bytes(_dlrm_like())
    m = ex.model
    tables = synth.make_tables_np(m)
    for batch in (64, 256):
        feats = _features_for(ex, batch, seed=batch)
        want, _ = _expected(ex, batch, feats, tables)
        got, _, status, _ = run_gpu(m, pack_request(m, batch, feats), tables, want_ids=False)
        assert status == 0
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_frozen_graph_tables_are_const_nodes():
    """convert_variables_to_constants turns every VariableV2 into a Const: the tables are found all the same
    (LookupOptimizer::Match accepts Const weights, lookup_optimizer.cc:64-65), their values come out as views of
    the serialized graph, and the dense layer's rank-2 Const kernel is not taken for a table."""
    b = Builder(frozen=True)
    outs = [b.bucketized("bucketize0", np.arange(0, 500, 5), 8),
            b.hashed_dense("categ_hashbucket0", 1000, 8, string=True),
            b.hashed_sparse("categ_hashbucket_sparse0", 1000, 16, rank=2, combiner="sum")]
    data = b.finish(outs)
    ex = G.extract_from_bytes(data)
    assert ex.skipped == [] and [c.name for c in ex.model.columns] == ["bucketize0", "categ_hashbucket0",
                                                                        "categ_hashbucket_sparse0/values"]
    assert [t.shape for t in ex.table_values] == [(101, 8), (1000, 8), (1000, 16)]
    assert all(t.dtype == np.float32 for t in ex.table_values)
    g = G.parse_graphdef(data)
    assert "dense/kernel" not in [t.name for t in G.find_tables(g)]
    # the extracted tables drive the oracle directly
    m, batch = ex.model, 32
    feats = _features_for(ex, batch)
    out, _ = _expected(ex, batch, feats, [np.array(t) for t in ex.table_values])
    c = m.columns[0]
    ids = np.searchsorted(np.arange(0, 500, 5, dtype=np.float32), feats[c.key].values, side="right")
    np.testing.assert_array_equal(out[:, :8], ex.table_values[0][ids])


def test_columns_sharing_one_table():
    """shared_embedding_columns / BASELINE config 5: several lookup chains read one variable - one column each,
    one table index for all of them."""
    b = Builder()
    o0 = b.hashed_sparse("user_clicks", 5000, 16, combiner="mean")
    read = b.last_read
    o1 = b.hashed_sparse("user_views", 5000, 16, combiner="sum", shared_read=read)
    o2 = b.bucketized("price", [1.0, 10.0, 100.0], 8)
    ex = G.extract_from_bytes(b.finish([o0, o1, o2]))
    assert ex.skipped == []
    m = ex.model
    assert [c.name for c in m.columns] == ["user_clicks/values", "user_views/values", "price"]
    assert [c.table for c in m.columns] == [0, 0, 1] and m.table_shapes == [(5000, 16), (4, 8)]
    assert [c.combiner for c in m.columns[:2]] == [L.COMBINE_MEAN, L.COMBINE_SUM]
    assert ex.bindings[0].table_variable == ex.bindings[1].table_variable
    # runs: both columns gather from the same table
    batch = 16
    feats = _features_for(ex, batch)
    tables = synth.make_tables_np(m)
    out, _ = run_oracle(m, pack_request(m, batch, feats), tables)
    assert out.shape == (batch, 16 + 16 + 8) and np.isfinite(out).all()
