"""SURVEY 8(f) N3 - tf.strings.split on the device. A column fed ONE ';'-joined string per row
(examples/python/dlrm.py:49-53) through RECOM_SEG_SPLIT_STR must equal the oracle on the same
column fed the tokens that the numpy restatement of StringSplit (oracle.string_split,
string_split_op.cc:37-99) produces."""
import numpy as np
import pytest

from oracle import oracle as O
from recom_b200 import _lib as L
from recom_b200 import consts, synth
from recom_b200.schema import Column, Feature, Model, Strings, pack_request
from tests.helpers import run_gpu, run_oracle

pytestmark = pytest.mark.gpu

ALPHA = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"


def _joined(rng, batch, delim=b";", max_tokens=10, empty_row=0.1, empty_token=0.0, tok_len=(8, 9)):
    rows = []
    for _ in range(batch):
        if rng.random() < empty_row:
            rows.append(b"")
            continue
        toks = []
        for _ in range(int(rng.integers(1, max_tokens + 1))):
            if rng.random() < empty_token:
                toks.append(b"")
            else:
                n = int(rng.integers(tok_len[0], tok_len[1]))
                toks.append(bytes(rng.choice(list(ALPHA), size=n).tolist()))
        rows.append(delim[:1].join(toks))
    return rows


def _models(delim, skip_empty):
    def mk(split):
        cols = [
            Column("a", L.LOOKUP_SEGMENT, L.XFORM_HASH_STR, L.STR, 0, 1000, 8, combiner=L.COMBINE_MEAN,
                   num_buckets=1000,
                   seg_format=L.SEG_SPLIT_STR if split else L.SEG_ROW_SPLITS_I32,
                   seg_stride=consts.split_spec(delim, skip_empty) if split else 2),
            Column("dense", L.LOOKUP_DENSE, L.XFORM_HASH_INT, L.I32, 1, 100, 8, num_buckets=100),
            # a second column over the same split input (one split + hash for both)
            Column("b", L.LOOKUP_SEGMENT, L.XFORM_HASH_STR, L.STR, 2, 1000, 32, combiner=L.COMBINE_SQRTN,
                   num_buckets=1000, input_key="a",
                   seg_format=L.SEG_SPLIT_STR if split else L.SEG_ROW_SPLITS_I32,
                   seg_stride=consts.split_spec(delim, skip_empty) if split else 2),
        ]
        return Model(cols, [(1000, 8), (100, 8), (1000, 32)], name="split" if split else "presplit").finalize()
    return mk(True), mk(False)


@pytest.mark.parametrize("batch,delim,skip_empty,empty_token,tok_len", [
    (2048, b";", True, 0.0, (8, 9)),      # the reference's model E inputs
    (333, b";", True, 0.2, (1, 20)),      # empty tokens dropped, ragged lengths, tokens beyond 16 bytes
    (100, b";", False, 0.3, (1, 12)),     # empty tokens kept (hash of the empty string)
    (77, b";,", True, 0.1, (2, 9)),       # delimiter set
    (50, b"", True, 0.0, (1, 6)),         # empty delimiter: character by character
])
def test_device_split_matches_string_split_op(batch, delim, skip_empty, empty_token, tok_len):
    rng = np.random.default_rng(batch)
    rows = _joined(rng, batch, delim if delim else b";", empty_token=empty_token, tok_len=tok_len)
    if len(delim) > 1:  # mix the delimiters of the set
        rows = [r.replace(b";", b",", 1) for r in rows]
    split_m, pre_m = _models(delim, skip_empty)
    tables = synth.make_tables_np(split_m, seed=1)
    ints = rng.integers(0, 1000, size=batch).astype(np.int32)
    got, _, status, launches = run_gpu(split_m, pack_request(split_m, batch, {
        "a": Feature(Strings.from_list(rows)), "dense": Feature(ints)}), tables, want_ids=False)
    assert status == 0 and launches == 2  # the split kernel + the fused kernel
    tokens, rs = O.string_split(rows, delim, skip_empty)
    ref, _ = run_oracle(pre_m, pack_request(pre_m, batch, {
        "a": Feature(Strings.from_list(tokens), row_splits=rs), "dense": Feature(ints)}), tables)
    err = np.abs(got - ref)
    assert np.all(err <= 1e-5 * np.maximum(np.abs(ref), 1.0)), float(err.max())
