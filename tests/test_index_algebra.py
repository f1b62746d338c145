"""the index algebra the reference inlines into its loads -
SparseReshape in front of the row id (cuda_emitter.cc:1874-1922) and Sum(axis=1) over a gathered
[batch, T, D] tensor (EmitBatchColReduction, cuda_emitter.cc:1180-1244) - as two segment formats
(RECOM_SEG_COO_RESHAPE_I64, RECOM_SEG_FIXED).

CPU: the oracle's restatement against numpy's own ravel / unravel (what TF's SparseReshape
kernel does) and against table[ids].sum(1). GPU: the CUDA path (row_splits built in the workspace
by coo_to_row_splits_kernel) against the oracle, and against the same lookup fed with host-made
row_splits."""
import numpy as np
import pytest

from recom_b200 import consts as L
from recom_b200.schema import Column, Feature, Model, pack_request
from tests.helpers import run_oracle


def _reshape_model(rank, has_new_shape, combiner=L.COMBINE_MEAN, dim=8, rows=500, n_cols=2):
    cols = [Column(f"r{i}", L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I64, i, rows, dim, combiner=combiner,
                   seg_format=L.SEG_COO_RESHAPE_I64, seg_stride=L.reshape_spec(rank, has_new_shape))
            for i in range(n_cols)]
    return Model(cols, [(rows, dim)] * n_cols, name="reshape").finalize()


def _sparse_input(rng, shape, density=0.3):
    """A row-major sorted SparseTensor of `shape`: (indices int64[nnz, rank], nnz)."""
    total = int(np.prod(shape))
    keep = np.flatnonzero(rng.random(total) < density)
    idx = np.stack(np.unravel_index(keep, shape), axis=1).astype(np.int64)
    return idx


def _tf_sparse_reshape_rows(idx, shape, new_shape):
    """numpy restatement of TF SparseReshape followed by [:, 0]."""
    new_shape = list(new_shape)
    total = int(np.prod(shape))
    if -1 in new_shape:
        k = new_shape.index(-1)
        known = int(np.prod([d for d in new_shape if d != -1]))
        new_shape[k] = total // known
    lin = np.ravel_multi_index(tuple(idx.T), shape) if idx.size else np.zeros(0, np.int64)
    return (np.stack(np.unravel_index(lin, new_shape), axis=1)[:, 0] if idx.size else lin), new_shape


def _reshape_case(seed, shape, new_shape):
    rng = np.random.default_rng(seed)
    rank = len(shape)
    m = _reshape_model(rank, new_shape is not None)
    feats, rows_ref = {}, {}
    batch = None
    for c in m.columns:
        idx = _sparse_input(rng, shape)
        ns = new_shape if new_shape is not None else [int(np.prod(shape[:-1])), shape[-1]]
        rows, full = _tf_sparse_reshape_rows(idx, shape, ns)
        batch = full[0]
        ids = rng.integers(0, c.table_rows, size=idx.shape[0]).astype(np.int64)
        feats[c.key] = Feature(ids, coo=idx, dense_shape=np.array(shape, np.int64),
                               new_shape=None if new_shape is None else np.array(new_shape, np.int64))
        rows_ref[c.key] = (rows, ids)
    return m, batch, feats, rows_ref


def _tables(m, seed=1):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((r, d)).astype(np.float32) for r, d in m.table_shapes]


def _expect(m, batch, rows_ref, tables):
    out = np.zeros((batch, m.out_width), np.float32)
    for c in m.columns:
        rows, ids = rows_ref[c.key]
        for b in range(batch):
            sel = ids[rows == b]
            if sel.size == 0:
                continue
            acc = np.zeros(c.dim, np.float32)
            for i in sel:  # sequential fp32 sum, the oracle's order
                acc += tables[c.table][i]
            if c.combiner == L.COMBINE_MEAN:
                acc = acc * np.float32(1.0 / sel.size) if sel.size < 10 else acc / np.float32(sel.size)
            out[b, c.out_offset:c.out_offset + c.dim] = acc
    return out


CASES = [((6, 5), None), ((4, 3, 5), None), ((4, 3, 5), [-1, 5]), ((4, 3, 5), [6, 10]),
         ((2, 3, 2, 4), [6, -1, 4]), ((3, 4, 2), [12, 2])]


@pytest.mark.parametrize("shape,new_shape", CASES)
def test_oracle_sparse_reshape_rows(shape, new_shape):
    m, batch, feats, rows_ref = _reshape_case(3, shape, new_shape)
    tables = _tables(m)
    req = pack_request(m, batch, feats)
    got, _ = run_oracle(m, req, tables)
    np.testing.assert_allclose(got, _expect(m, batch, rows_ref, tables), rtol=1e-6, atol=1e-6)


def _fixed_model(T, static, dtype=L.I64, dim=16, rows=300):
    cols = [Column("seq0", L.LOOKUP_SEGMENT, L.XFORM_NONE, dtype, 0, rows, dim, combiner=L.COMBINE_SUM,
                   seg_format=L.SEG_FIXED, seg_stride=T if static else 0),
            Column("seq1", L.LOOKUP_SEGMENT, L.XFORM_HASH_INT, L.I32, 1, rows, 8, combiner=L.COMBINE_MEAN,
                   num_buckets=rows, seg_format=L.SEG_FIXED, seg_stride=T if static else 0)]
    return Model(cols, [(rows, dim), (rows, 8)], name="fixed").finalize()


@pytest.mark.parametrize("static", [True, False])
def test_oracle_batch_col_reduction(static):
    T, batch = 7, 33
    m = _fixed_model(T, static)
    rng = np.random.default_rng(0)
    ids0 = rng.integers(0, 300, size=(batch, T)).astype(np.int64)
    raw1 = rng.integers(0, 100, size=(batch, T)).astype(np.int32)
    feats = {"seq0": Feature(ids0.reshape(-1)), "seq1": Feature(raw1.reshape(-1))}
    tables = _tables(m)
    req = pack_request(m, batch, feats)
    got, ids = run_oracle(m, req, tables)
    # Sum(axis=1) of the gathered [batch, T, D] tensor (cuda_emitter.cc:1212-1240)
    np.testing.assert_allclose(got[:, :16], tables[0][ids0].sum(axis=1), rtol=1e-5, atol=1e-5)
    hashed = ids[batch * T:].reshape(batch, T)
    np.testing.assert_allclose(got[:, 16:], tables[1][hashed].mean(axis=1), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,new_shape", CASES + [((64, 7, 9), None), ((300, 4, 6), [-1, 6])])
def test_gpu_sparse_reshape(shape, new_shape):
    from tests.helpers import assert_parity

    m, batch, feats, rows_ref = _reshape_case(11, shape, new_shape)
    tables = _tables(m)
    _, got, _ = assert_parity(m, batch, feats, tables)
    np.testing.assert_allclose(got, _expect(m, batch, rows_ref, tables), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("static", [True, False])
@pytest.mark.parametrize("T,batch", [(1, 5), (7, 33), (50, 257), (130, 40)])
def test_gpu_batch_col_reduction(static, T, batch):
    from tests.helpers import assert_parity

    m = _fixed_model(T, static)
    rng = np.random.default_rng(T)
    ids0 = rng.integers(0, 300, size=(batch, T)).astype(np.int64)
    raw1 = rng.integers(0, 100, size=(batch, T)).astype(np.int32)
    feats = {"seq0": Feature(ids0.reshape(-1)), "seq1": Feature(raw1.reshape(-1))}
    tables = _tables(m)
    _, got, _ = assert_parity(m, batch, feats, tables)
    np.testing.assert_allclose(got[:, :16], tables[0][ids0].sum(axis=1), rtol=1e-5, atol=2e-5)
