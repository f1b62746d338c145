"""The CPU oracle against TensorFlow's own known-answer vectors (tests/golden/tf_golden.json,
each entry cites the TF test it was transcribed from). This is what pins the oracle before it
is trusted as the checker of the CUDA path. CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tf_golden.json")))
SUM, MEAN, SQRTN = 0, 1, 2
DENSE, GATHER_SCATTER, SEGMENT = 0, 1, 2


def test_fingerprint64_text_vectors():
    for c in G["fingerprint64"]["cases"]:
        assert O.fingerprint64(c["text"].encode()) == int(c["u64"]), c


def test_fingerprint_op_byte_vectors():
    # fingerprint_op_test.cc: regimes len 0, 1-16, 17-32 and >64
    for c in G["fingerprint_op"]["cases"]:
        s = bytes((c["start"] + i) % 256 for i in range(c["len"]))
        got = O.fingerprint64(s).to_bytes(8, "little").hex()
        assert got == c["le_bytes"], c


def test_fingerprint64_batch_matches_scalar():
    rng = np.random.default_rng(0)
    strs = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in
            list(range(0, 140)) + [255, 256, 257, 1000]]
    off = np.zeros(len(strs) + 1, np.int32)
    off[1:] = np.cumsum([len(s) for s in strs])
    data = np.frombuffer(b"".join(strs), np.uint8)
    got = O.fingerprint64_batch(data, off)
    want = np.array([O.fingerprint64(s) for s in strs], np.uint64)
    np.testing.assert_array_equal(got, want)


def test_string_to_hash_bucket_fast_vectors():
    for c in G["string_to_hash_bucket_fast"]["cases"]:
        bs = [s.encode() for s in c["strings"]]
        off = np.zeros(len(bs) + 1, np.int32)
        off[1:] = np.cumsum([len(b) for b in bs])
        got = O.string_to_hash_bucket_fast(np.frombuffer(b"".join(bs), np.uint8), off, c["num_buckets"])
        assert got.tolist() == c["expect"], c


def test_hash_int_is_as_string_then_hash():
    # feature_column_v2.py:3414-3419: int inputs are rendered with as_string first
    vals = np.array([0, 7, 10, 99, 100, 12345, -1, -987654321, 2**31 - 1, -2**31], np.int64)
    want = [O.fingerprint64(str(int(v)).encode()) % 1000 for v in vals]
    assert O.hash_int(vals, 1000).tolist() == want
    v32 = vals.astype(np.int32)
    assert O.hash_int(v32, 1000).tolist() == want
    big = np.array([2**63 - 1, -2**63, 10**18], np.int64)
    assert O.hash_int(big, 77).tolist() == [O.fingerprint64(str(int(v)).encode()) % 77 for v in big]


def test_bucketize_vectors():
    for c in G["bucketize"]["cases"]:
        v = np.array(c["values"], dtype=c["dtype"])
        got = O.bucketize(v, np.array(c["boundaries"], np.float32))
        assert got.tolist() == c["expect"], c


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
def test_bucketize_is_upper_bound(dtype):
    # bucketize_op.cc:42-44: std::upper_bound over float boundaries
    rng = np.random.default_rng(1)
    b = np.sort(rng.integers(-50, 50, size=37)).astype(np.float32)  # with duplicates
    v = rng.integers(-60, 60, size=500).astype(dtype)
    want = np.searchsorted(b.astype(np.float64), v.astype(np.float64), side="right")
    np.testing.assert_array_equal(O.bucketize(v, b), want)


def test_bucketize_nan_and_inf():
    b = np.array([0.0, 1.0, 2.0], np.float32)
    v = np.array([np.nan, np.inf, -np.inf], np.float32)
    # upper_bound with `value < boundary`: NaN compares false everywhere -> past the end
    assert O.bucketize(v, b).tolist() == [3, 3, 0]


def test_sparse_fill_empty_rows_vector():
    g = G["sparse_fill_empty_rows"]
    oi, ov, empty, rev = O.sparse_fill_empty_rows(np.array(g["indices"]), np.array(g["values"]),
                                                  g["dense_shape"][0], g["default"])
    assert oi.tolist() == g["out_indices"]
    assert ov.tolist() == g["out_values"]
    assert empty.astype(int).tolist() == g["empty_row_indicator"]
    assert rev.tolist() == [0, 1, 2, 3, 5, 6]


def _np_segment(table, idx, seg, rows, combiner):
    out = np.zeros((rows, table.shape[1]), np.float64)
    cnt = np.zeros(rows)
    for i, s in zip(idx, seg):
        out[s] += table[i]
        cnt[s] += 1
    if combiner == MEAN:
        out /= np.maximum(cnt, 1)[:, None]
    if combiner == SQRTN:
        out /= np.sqrt(np.maximum(cnt, 1))[:, None]
    return out


@pytest.mark.parametrize("combiner", [SUM, MEAN, SQRTN])
def test_sparse_segment_vectors(combiner):
    g = G["sparse_segment"]
    r, d = g["table_shape"]
    table = np.arange(1, r * d + 1, dtype=np.float32).reshape(r, d)  # _input()
    for c in g["cases"]:
        idx, seg, ns = np.array(c["indices"], np.int64), np.array(c["segment_ids"], np.int64), c["num_segments"]
        rows = ns if ns >= 0 else (int(seg[-1]) + 1 if seg.size else 0)
        got = O.sparse_segment_reduce(table, idx, seg, ns, combiner)
        np.testing.assert_allclose(got, _np_segment(table, idx, seg, rows, combiner), rtol=1e-6)


def test_sparse_segment_invalid_inputs():
    table = np.ones((10, 4), np.float32)
    with pytest.raises(ValueError):  # segment ids not increasing (testSegmentsInvalid*)
        O.sparse_segment_reduce(table, np.array([1, 2, 3]), np.array([0, 2, 1]), -1, SUM)
    with pytest.raises(ValueError):  # index out of range
        O.sparse_segment_reduce(table, np.array([1, 10]), np.array([0, 1]), -1, SUM)
    with pytest.raises(ValueError):  # num_segments too small
        O.sparse_segment_reduce(table, np.array([1, 2]), np.array([0, 3]), 2, SUM)


def test_sparse_segment_long_segments_random():
    # exercises TF's 8-row block order and the n >= 10 divide (segment_reduction_ops_impl.h:619-770)
    rng = np.random.default_rng(2)
    table = rng.standard_normal((500, 12)).astype(np.float32)
    lens = np.array([1, 2, 7, 8, 9, 10, 11, 16, 17, 33, 100, 0, 3])
    seg = np.repeat(np.arange(lens.size), lens)
    idx = rng.integers(0, 500, size=seg.size)
    for comb in (SUM, MEAN, SQRTN):
        got = O.sparse_segment_reduce(table, idx, seg, lens.size, comb)
        np.testing.assert_allclose(got, _np_segment(table, idx, seg, lens.size, comb), rtol=2e-5, atol=2e-6)


def test_safe_embedding_lookup_sparse_no_weights_vector():
    g = G["safe_embedding_lookup_sparse"]
    rng = np.random.default_rng(3)
    w = rng.standard_normal((3, 4)).astype(np.float32)
    got = O.safe_embedding_lookup_sparse(w, np.array(g["indices"]), np.array(g["ids"]), g["dense_rows"], MEAN)
    rows = {"mean01": (w[0] + w[1]) / 2.0, "zero": np.zeros(4, np.float32), "row2": w[2]}
    want = np.stack([rows[k] for k in g["expect_rows"]])
    np.testing.assert_allclose(got, want, rtol=1e-6)


def test_embedding_column_vector():
    g = G["embedding_column"]
    table = np.array(g["table"], np.float32)
    got = O.safe_embedding_lookup_sparse(table, np.array(g["indices"]), np.array(g["ids"]), g["dense_rows"], MEAN)
    np.testing.assert_array_equal(got, np.array(g["expect"], np.float32))
    # the same column through RECom's SEGMENT rewrite (lookup_optimizer.cc:157-268)
    rows = np.array(g["indices"])[:, 0]
    got2 = O.recom_lookup(table, SEGMENT, MEAN, np.array(g["ids"]), rows, g["dense_rows"])
    np.testing.assert_array_equal(got2, got)


# --- RECom's three rewrites == the literal TF chain (unpinned in the reference) ---
def _coo(lengths, rank=2):
    rows = np.repeat(np.arange(lengths.size), lengths)
    coo = np.zeros((rows.size, rank), np.int64)
    coo[:, 0] = rows
    return coo, rows


@pytest.mark.parametrize("combiner", [SUM, MEAN, SQRTN])
@pytest.mark.parametrize("bf16", [False, True])
def test_segment_rewrite_equals_tf_chain(combiner, bf16):
    rng = np.random.default_rng(4 + combiner)
    table = rng.standard_normal((300, 16)).astype(np.float32)
    t = O.f32_to_bf16_bits(table) if bf16 else table
    lengths = rng.integers(0, 14, size=64)
    coo, rows = _coo(lengths)
    ids = rng.integers(0, 300, size=rows.size)
    a = O.safe_embedding_lookup_sparse(t, coo, ids, 64, combiner)
    b = O.recom_lookup(t, SEGMENT, combiner, ids, rows, 64)
    np.testing.assert_array_equal(a, b)


def test_gather_scatter_rewrite_equals_tf_chain():
    rng = np.random.default_rng(5)
    table = rng.standard_normal((50, 8)).astype(np.float32)
    lengths = (rng.random(40) < 0.6).astype(np.int64)
    coo, rows = _coo(lengths, rank=1)
    ids = rng.integers(0, 50, size=rows.size)
    for comb in (SUM, MEAN):
        a = O.safe_embedding_lookup_sparse(table, coo, ids, 40, comb)
        b = O.recom_lookup(table, GATHER_SCATTER, SUM, ids, rows, 40)
        np.testing.assert_array_equal(a, b)


def test_dense_rewrite_equals_tf_chain():
    rng = np.random.default_rng(6)
    table = rng.standard_normal((50, 8)).astype(np.float32)
    coo, rows = _coo(np.ones(33, np.int64))
    ids = rng.integers(0, 50, size=33)
    a = O.safe_embedding_lookup_sparse(table, coo, ids, 33, MEAN)
    b = O.recom_lookup(table, DENSE, SUM, ids, None, 33)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(b, table[ids])


def test_bf16_round_trip_is_rne():
    x = np.array([1.0, 1.00390625, 1.005859375, -2.5, 3.0e38, 1e-40, 0.0], np.float32)
    bits = O.f32_to_bf16_bits(x)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(bits, want)
    np.testing.assert_array_equal(O.bf16_bits_to_f32(bits), torch.from_numpy(x).to(torch.bfloat16).float().numpy())
