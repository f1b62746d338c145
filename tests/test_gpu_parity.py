"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import synth
from tests.helpers import assert_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("batch", [1, 7, 256, 300])
def test_model_e_small(batch):
    m = synth.model_E_small()
    feats = synth.make_features(m, batch, seed=batch)
    assert_parity(m, batch, feats, synth.make_tables_np(m))


@pytest.mark.parametrize("fmt", [L.SEG_ROW_SPLITS_I32, L.SEG_COO_I64])
@pytest.mark.parametrize("fused", [True, False])
def test_model_e_small_formats(fmt, fused):
    m = synth.model_E_small(seg_format=fmt, fused_hash=fused)
    feats = synth.make_features(m, 129, seed=3, empty_frac=0.2)
    assert_parity(m, 129, feats, synth.make_tables_np(m))


def test_microbenchmark_100():
    m = synth.microbenchmark_model(100)
    batch = 256
    feats = synth.make_features(m, batch, seed=0)
    # microbenchmark.py:66 feeds randint(-1, 10000) cast to float
    rng = np.random.default_rng(0)
    for c in m.columns:
        feats[c.key].values = rng.integers(-1, 10000, size=batch).astype(np.float32)
    assert_parity(m, batch, feats, synth.make_tables_np(m))


@pytest.mark.parametrize("out_dtype", [L.F32, L.BF16])
def test_mixed_model(out_dtype):
    m = synth.mixed_model(120, seed=1, out_dtype=out_dtype, large_rows=1 << 15)
    feats = synth.make_features(m, 200, seed=5, dist="u50", empty_frac=0.1)
    assert_parity(m, 200, feats, synth.make_tables_np(m), check_ids=False)


def test_mixed_model_checked_ids():
    m = synth.mixed_model(60, seed=2, large_rows=1 << 14)
    feats = synth.make_features(m, 64, seed=6, dist="zipf")
    assert_parity(m, 64, feats, synth.make_tables_np(m), check_ids=True)


# ---- committed golden fixtures ---------------------------------------------------------
import os

from tests.golden.make_fixtures import CASES, build_case
from tests.helpers import run_gpu

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage_fixtures.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_against_committed_fixture(name):
    model, req, tables = build_case(name)
    got, ids, status, _ = run_gpu(model, req, tables)
    assert status == 0
    np.testing.assert_array_equal(ids, FX[f"{name}/ids"])  # bit exact
    ref = FX[f"{name}/out"]
    rtol = 1e-2 if model.out_dtype == L.BF16 else 1e-5
    # per-element bound against the summed magnitude (see helpers.assert_parity)
    from tests.helpers import run_oracle
    scale, _ = run_oracle(model, req, [np.abs(t) for t in tables])
    assert (np.abs(got - ref) <= rtol * scale + 1e-30).all()


# ---- edge cases ------------------------------------------------------------------------
def test_batch_zero_is_a_no_op():
    m = synth.model_E_small()
    assert_parity(m, 0, synth.make_features(m, 0, seed=0), synth.make_tables_np(m))


def test_all_rows_empty():
    m = synth.model_E_small()
    feats = synth.make_features(m, 40, seed=1, empty_frac=1.0)
    req, got, ref = assert_parity(m, 40, feats, synth.make_tables_np(m))
    for c in m.columns:
        if c.lookup != L.LOOKUP_DENSE:
            assert (got[:, c.out_offset:c.out_offset + c.dim] == 0).all()


@pytest.mark.parametrize("dist", ["zipf", "u50", "one"])
def test_ragged_lengths(dist):
    m = synth.mixed_model(50, seed=4, large_rows=1 << 14, frac_multihot=0.5)
    feats = synth.make_features(m, 96, seed=8, dist=dist, empty_frac=0.15)
    assert_parity(m, 96, feats, synth.make_tables_np(m))


def test_one_very_long_row():
    # a 5000-id row next to empty rows: one segment far longer than a tile
    m = synth.dlrm_model(2, 0, 0, 2, 1, large_rows=1 << 12, fused_hash=False)
    batch = 9
    feats = synth.make_features(m, batch, seed=3)
    rng = np.random.default_rng(0)
    for c in m.columns:
        if c.lookup == L.LOOKUP_SEGMENT:
            lengths = np.zeros(batch, np.int64)
            lengths[4] = 5000
            lengths[8] = 3
            rs, coo = synth.splits_and_coo(lengths)
            feats[c.key].row_splits, feats[c.key].coo = rs, coo
            feats[c.key].values = rng.integers(0, c.table_rows, size=int(rs[-1])).astype(np.int64)
    assert_parity(m, batch, feats, synth.make_tables_np(m))


def test_unaligned_dims_fall_back_to_scalar_chunks():
    # dims that are not multiples of 16 bytes, and a row length that is not either
    from recom_b200.schema import Column, Model
    cols, tables = [], []
    for i, d in enumerate([3, 5, 8, 1, 7, 12, 2]):
        b = np.arange(0, 50, 5, dtype=np.float32)
        lookup = L.LOOKUP_DENSE if i % 2 == 0 else L.LOOKUP_SEGMENT
        if lookup == L.LOOKUP_DENSE:
            cols.append(Column(f"c{i}", lookup, L.XFORM_BUCKETIZE, L.F32, i, 11, d, boundaries=b))
            tables.append((11, d))
        else:
            cols.append(Column(f"c{i}", lookup, L.XFORM_NONE, L.I64, i, 200, d, combiner=L.COMBINE_MEAN,
                               seg_format=L.SEG_ROW_SPLITS_I32))
            tables.append((200, d))
    for dt in (L.F32, L.BF16):
        m = Model(cols, tables, out_dtype=dt).finalize()
        feats = synth.make_features(m, 77, seed=9, empty_frac=0.2)
        assert_parity(m, 77, feats, synth.make_tables_np(m))


def test_wide_rows():
    from recom_b200.schema import Column, Model
    cols = [Column("w0", L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I32, 0, 300, 256, combiner=L.COMBINE_SQRTN,
                   seg_format=L.SEG_ROW_SPLITS_I32),
            Column("w1", L.LOOKUP_DENSE, L.XFORM_NONE, L.I64, 1, 64, 192)]
    m = Model(cols, [(300, 256), (64, 192)]).finalize()
    feats = synth.make_features(m, 50, seed=10, dist="u50")
    assert_parity(m, 50, feats, synth.make_tables_np(m))


def test_production_model_shared_tables_and_inputs():
    m = synth.production_model(300, share=4, seed=1)
    feats = synth.make_features(m, 64, seed=2)
    assert_parity(m, 64, feats, synth.make_tables_np(m))


def test_out_of_range_id_is_flagged_not_fatal():
    m = synth.dlrm_model(2, 0, 0, 1, 0, fused_hash=False)
    feats = synth.make_features(m, 16, seed=0)
    for c in m.columns:
        if c.lookup == L.LOOKUP_SEGMENT:
            feats[c.key].values[3] = c.table_rows + 5
    from recom_b200.schema import pack_request
    req = pack_request(m, 16, feats)
    got, ids, status, _ = run_gpu(m, req, synth.make_tables_np(m), check_ids=True)
    assert status == L.ERR_ID_OUT_OF_RANGE
    assert np.isfinite(got).all()


def test_workspace_too_small_is_an_error():
    import torch
    from recom_b200 import runtime as R
    from recom_b200.schema import pack_request
    m = synth.model_E_small()
    req = pack_request(m, 64, synth.make_features(m, 64, seed=0))
    plan = R.Plan(m)
    tables = R.tables_from_numpy(m, synth.make_tables_np(m))
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    with pytest.raises(L.RecomError) as e:
        plan.run(req, R.upload_blob(req), R.table_pointers(tables), plan.alloc_out(64), ws)
    assert e.value.status == L.ERR_WORKSPACE_TOO_SMALL


def test_plan_serialize_round_trip():
    from recom_b200 import runtime as R
    from recom_b200.schema import pack_request
    m = synth.model_E_small()
    tables = synth.make_tables_np(m)
    req = pack_request(m, 48, synth.make_features(m, 48, seed=4))
    p1 = R.Plan(m)
    p2 = R.Plan.from_blob(p1.serialize(), m)
    a, ia, _, _ = run_gpu(m, req, tables, plan=p1)
    b, ib, _, _ = run_gpu(m, req, tables, plan=p2)
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(ia, ib)


def test_reentrant_runs_on_two_streams():
    import torch
    from recom_b200 import runtime as R
    from recom_b200.schema import pack_request
    m = synth.model_E_small()
    tables_np = synth.make_tables_np(m)
    plan = R.Plan(m)
    tables = R.tables_from_numpy(m, tables_np)
    ptrs = R.table_pointers(tables)
    streams = [torch.cuda.Stream() for _ in range(2)]
    reqs = [pack_request(m, 200, synth.make_features(m, 200, seed=20 + i)) for i in range(2)]
    blobs = [R.upload_blob(r) for r in reqs]
    outs = [plan.alloc_out(200) for _ in reqs]
    wss = [plan.alloc_workspace(200, r.sparse_ids) for r in reqs]
    torch.cuda.synchronize()
    for it in range(20):
        for i in range(2):
            plan.run(reqs[i], blobs[i], ptrs, outs[i], wss[i], stream=streams[i])
    torch.cuda.synchronize()
    from tests.helpers import run_oracle
    for i in range(2):
        ref, _ = run_oracle(m, reqs[i], tables_np)
        np.testing.assert_allclose(outs[i].cpu().numpy(), ref, rtol=1e-5, atol=1e-6)


# ---- device hashing against the oracle, every length regime ------------------------------
def test_device_fingerprint64_all_regimes():
    import torch
    from oracle import oracle as O
    from recom_b200 import runtime as R
    rng = np.random.default_rng(0)
    lens = list(range(0, 200)) + [255, 256, 257, 511, 1000, 2520]
    strs = [bytes(rng.integers(0, 256, size=n, dtype=np.uint8)) for n in lens for _ in range(3)]
    off = np.zeros(len(strs) + 1, np.int32)
    off[1:] = np.cumsum([len(s) for s in strs])
    data = np.frombuffer(b"".join(strs), np.uint8).copy()
    pad = np.zeros((data.size + 15) // 16 * 16 + 16, np.uint8)
    pad[:data.size] = data
    got = R.fingerprint64(torch.from_numpy(pad).cuda(), torch.from_numpy(off).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got.view(np.uint64), O.fingerprint64_batch(data, off))
    for nb in (1, 10, 10000, 1 << 23, 2**31 - 1):
        b = R.string_to_hash_bucket_fast(torch.from_numpy(pad).cuda(), torch.from_numpy(off).cuda(), nb)
        np.testing.assert_array_equal(b.cpu().numpy(), O.string_to_hash_bucket_fast(data, off, nb))


def test_device_golden_hash_vectors():
    import json
    import torch
    from recom_b200 import runtime as R
    from recom_b200.schema import Strings
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tf_golden.json")))
    for c in g["string_to_hash_bucket_fast"]["cases"]:
        s = Strings.from_list(c["strings"])
        pad = np.zeros(s.data.size + 32, np.uint8)
        pad[:s.data.size] = s.data
        got = R.string_to_hash_bucket_fast(torch.from_numpy(pad).cuda(), torch.from_numpy(s.offsets).cuda(),
                                           c["num_buckets"])
        assert got.cpu().tolist() == c["expect"]


def test_hash_int_extremes():
    from recom_b200.schema import Column, Model
    cols = [Column("a", L.LOOKUP_DENSE, L.XFORM_HASH_INT, L.I64, 0, 977, 8, num_buckets=977),
            Column("b", L.LOOKUP_DENSE, L.XFORM_HASH_INT, L.I32, 1, 1 << 20, 4, num_buckets=1 << 20)]
    m = Model(cols, [(977, 8), (1 << 20, 4)]).finalize()
    from recom_b200.schema import Feature
    v64 = np.array([0, 1, 9, 10, 99999999, 100000000, -1, -10, 2**63 - 1, -2**63, 10**18, -(10**18),
                    123456789012, 42, 7, 99], np.int64)
    v32 = np.array([0, 5, 10, 99999999, 100000000, 2**31 - 1, -2**31, -1, 12345, 77, 3, 2, 1, 100, 1000, 65536],
                   np.int32)
    feats = {"a": Feature(v64), "b": Feature(v32)}
    assert_parity(m, 16, feats, synth.make_tables_np(m))


def test_segment_ids_to_row_splits():
    import torch
    from recom_b200 import runtime as R
    rng = np.random.default_rng(5)
    for batch, stride in ((1, 1), (100, 2), (4096, 3)):
        lengths = synth.row_lengths(rng, batch, "u10", empty_frac=0.3)
        rs, coo = synth.splits_and_coo(lengths, stride)
        got = R.segment_ids_to_row_splits(torch.from_numpy(coo).cuda(), batch).cpu().numpy()
        np.testing.assert_array_equal(got, rs)


# ---- full-size properties (BASELINE.json sizes; the oracle is too slow to run these often) ----
def _full_size_run(model, batch, seed, plan=None):
    import torch
    from recom_b200 import runtime as R
    from recom_b200.schema import pack_request
    plan = plan or R.Plan(model)
    tables = R.make_tables(model, seed=0)
    req = pack_request(model, batch, synth.make_features(model, batch, seed=seed))
    out = plan.alloc_out(batch)
    out.fill_(float("nan"))
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    ids = torch.empty(req.total_ids, dtype=torch.int32, device="cuda")
    plan.run(req, R.upload_blob(req), R.table_pointers(tables), out, ws, ids_out=ids)
    torch.cuda.synchronize()
    return plan, tables, req, out, ids


def test_model_e_full_size_properties():
    import torch
    m = synth.model_E()
    B = 2048
    plan, tables, req, out, ids = _full_size_run(m, B, seed=1)
    assert not torch.isnan(out).any()
    # property 1: every DENSE column slice equals table[ids] exactly (pure copy)
    for ci in (0, 1, 499, 998, 999):
        c = m.columns[ci]
        if c.lookup != L.LOOKUP_DENSE:
            continue
        cid = ids[int(req.ids_base[ci]):int(req.ids_base[ci + 1])].long()
        assert int(cid.min()) >= 0 and int(cid.max()) < c.table_rows
        assert torch.equal(out[:, c.out_offset:c.out_offset + c.dim], tables[c.table][cid])
    # property 2: multi-hot sum columns equal an index_add of the gathered rows (fp32 tolerance)
    for ci, c in enumerate(m.columns):
        if c.lookup != L.LOOKUP_SEGMENT:
            continue
        cid = ids[int(req.ids_base[ci]):int(req.ids_base[ci + 1])].long()
        rs = torch.from_numpy(req.blob[int(req.slot_offsets[c.seg_slot]):][:4 * (B + 1)].view(np.int32).copy())
        seg = torch.repeat_interleave(torch.arange(B), (rs[1:] - rs[:-1]).long()).cuda()
        want = torch.zeros(B, c.dim, device="cuda").index_add_(0, seg, tables[c.table][cid])
        torch.testing.assert_close(out[:, c.out_offset:c.out_offset + c.dim], want, rtol=1e-5, atol=1e-6)
    # property 3: idempotence, the same request twice gives identical bits
    out2 = out.clone()
    plan2, _, _, out3, _ = _full_size_run(m, B, seed=1, plan=plan)
    assert torch.equal(out2, out3)


def test_model_e_rows_are_batch_independent():
    # linearity over the batch: running rows [0,B) equals running two halves separately
    import torch
    from recom_b200 import runtime as R
    from recom_b200.schema import Feature, Strings, pack_request
    m = synth.model_E_small()
    B = 512
    feats = synth.make_features(m, B, seed=3)
    tables_np = synth.make_tables_np(m)
    full, _, _, _ = run_gpu(m, pack_request(m, B, feats), tables_np)

    def half(f, lo, hi):
        if f.row_splits is None:
            if isinstance(f.values, Strings):
                o = f.values.offsets
                return Feature(Strings((o[lo:hi + 1] - o[lo]).astype(np.int32), f.values.data[o[lo]:o[hi]]))
            return Feature(f.values[lo:hi])
        rs = f.row_splits
        e0, e1 = int(rs[lo]), int(rs[hi])
        if isinstance(f.values, Strings):
            o = f.values.offsets
            vals = Strings((o[e0:e1 + 1] - o[e0]).astype(np.int32), f.values.data[o[e0]:o[e1]])
        else:
            vals = f.values[e0:e1]
        coo = f.coo[e0:e1].copy()
        coo[:, 0] -= lo
        return Feature(vals, row_splits=(rs[lo:hi + 1] - rs[lo]).astype(np.int32), coo=coo)

    parts = []
    for lo, hi in ((0, 200), (200, 512)):
        fh = {k: half(f, lo, hi) for k, f in feats.items()}
        got, _, _, _ = run_gpu(m, pack_request(m, hi - lo, fh), tables_np)
        parts.append(got)
    np.testing.assert_array_equal(full, np.concatenate(parts))
