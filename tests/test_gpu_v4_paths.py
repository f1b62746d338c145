"""GPU parity for the paths kernel v4 added: the bucketize grid index (fast groups), the tile
queue with its self-resetting claim counters, and the content-keyed argument-block cache.
Everything goes through the C ABI and is checked against the CPU oracle."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import synth
from recom_b200.schema import Column, Feature, Model, pack_request
from tests.helpers import assert_parity, run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def _boundary_sets():
    f = np.float32
    tiny = np.float32(1e-42)  # denormal
    return {
        "uniform": np.arange(0, 500, 5, dtype=f),                      # dlrm.py:126
        "duplicates": np.sort(np.repeat(np.array([0, 1, 1.5, 7, 7, 9, 100, 1e3], f), 3)),
        "clustered": np.concatenate([np.linspace(0, 1e-3, 40, dtype=f), np.array([1e6, 2e6], f)]),
        "huge_range": np.array([-3e38, -1e30, -1, 0, 1, 2, 3, 5, 1e30, 3e38], f),
        "denormal": np.array([k * tiny for k in range(1, 13)], f),
        "negative": np.linspace(-1000, -1, 64, dtype=f),
        "three": np.array([-1, 0, 1], f),                               # below the grid minimum
        "with_inf": np.array([-np.inf, -1, 0, 1, 2, 3, 4, 5, 6, np.inf], f),  # not grid-able
        "constant": np.full(9, 2.5, f),
        "many": np.sort(np.random.default_rng(3).normal(0, 50, 700).astype(f)),
    }


def _edge_values(b: np.ndarray, n: int, rng) -> np.ndarray:
    """n float32 values that stress upper_bound over b: the boundaries themselves, their float
    neighbours, zeros of both signs, infinities, NaN, far outliers, and random fill."""
    f = np.float32
    fin = b[np.isfinite(b)]
    special = [np.nan, np.inf, -np.inf, 0.0, -0.0, 3.4e38, -3.4e38, 1e-45, -1e-45]
    vals = list(special) + list(b) + list(np.nextafter(b, f(np.inf))) + list(np.nextafter(b, f(-np.inf)))
    lo, hi = (float(fin.min()), float(fin.max())) if fin.size else (-1.0, 1.0)
    span = max(hi - lo, 1e-30)
    vals += list(rng.uniform(lo - 0.1 * span, hi + 0.1 * span, size=max(n, 64)))
    with np.errstate(over="ignore"):
        vals = np.array(vals, dtype=f)
    rng.shuffle(vals)
    return np.resize(vals, n).astype(f)


@pytest.mark.parametrize("batch", [97, 1000])
def test_bucketize_grid_is_exact_on_edge_values(batch):
    """Fast groups (8 float32 bucketize columns): groups sharing one boundary set and groups
    with eight different sets; ids must equal std::upper_bound bit for bit (bucketize_op.cc:36-48)."""
    sets = _boundary_sets()
    names = sorted(sets)
    cols, tables = [], []
    plan = []
    for g in range(5):      # groups 0-4: one set per group (uniform groups)
        plan += [names[g]] * 8
    for k in range(16):     # groups 5-6: a different set per column
        plan.append(names[k % len(names)])
    for i, nm in enumerate(plan):
        b = sets[nm]
        cols.append(Column(f"c{i}", L.LOOKUP_DENSE, L.XFORM_BUCKETIZE, L.F32, i, len(b) + 1, 8, boundaries=b))
        tables.append((len(b) + 1, 8))
    m = Model(cols, tables).finalize()
    rng = np.random.default_rng(batch)
    feats = {c.key: Feature(_edge_values(sets[plan[i]], batch, rng)) for i, c in enumerate(m.columns)}
    req, got, ref = assert_parity(m, batch, feats, synth.make_tables_np(m))
    # gathers of single rows are copies: bit identical, not just within tolerance
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("batch", [61, 600])
def test_bucketize_grid_on_queue_tiles_all_value_dtypes(batch):
    """Bucketize columns that do NOT form fast groups (mixed value dtypes, dims and a hashed
    neighbour) go through the queue's generic tiles, whose search reads the set's grid from the
    global pool (bucketize_grid): float32, int32 and int64 values (integers compare as floats,
    like TF's upper_bound over float boundaries), multi-hot bucketize, every edge set."""
    sets = _boundary_sets()
    names = sorted(sets)
    cols, tables, feats = [], [], {}
    rng = np.random.default_rng(batch)
    big = [0, 1, -1, 2**24, 2**24 + 1, -(2**24) - 1, 2**31 - 1, -(2**31), 16777217, 33554433]
    for i in range(3 * len(names)):
        b = sets[names[i % len(names)]]
        dt = (L.F32, L.I32, L.I64)[i // len(names)]
        dim = (8, 16, 32)[i % 3]
        cols.append(Column(f"c{i}", L.LOOKUP_DENSE, L.XFORM_BUCKETIZE, dt, i, len(b) + 1, dim, boundaries=b))
        tables.append((len(b) + 1, dim))
        v = _edge_values(b, batch, rng)
        if dt == L.F32:
            vals = v
        else:
            with np.errstate(invalid="ignore"):
                iv = np.nan_to_num(v, nan=0.0, posinf=2e9, neginf=-2e9).clip(-2e9, 2e9)
            wide = np.int64 if dt == L.I64 else np.int32
            vals = np.rint(iv).astype(np.int64)
            vals[: len(big)] = big
            if dt == L.I64:
                vals[len(big): len(big) + 4] = [2**40 + 1, -(2**40) - 1, 2**62, -(2**62)]
            else:
                vals = vals.clip(-(2**31), 2**31 - 1)
            vals = vals.astype(wide)
        feats[cols[-1].key] = Feature(np.resize(vals, batch))
    n = len(cols)
    cols.append(Column("h", L.LOOKUP_DENSE, L.XFORM_HASH_INT, L.I32, n, 97, 8, num_buckets=97))
    tables.append((97, 8))
    feats[cols[-1].key] = Feature(rng.integers(0, 1000, batch).astype(np.int32))
    m = Model(cols, tables).finalize()
    req, got, ref = assert_parity(m, batch, feats, synth.make_tables_np(m))
    np.testing.assert_array_equal(got, ref)


def test_same_plan_many_requests_resets_its_queue_counters():
    """The queue's claim counters live in plan-owned slots that the last warp of a launch
    resets; results must not depend on how many launches (or which batch sizes) came before."""
    import torch
    from recom_b200 import runtime as R
    m = synth.model_E_small()
    tables_np = synth.make_tables_np(m)
    plan = R.Plan(m)
    tables = R.tables_from_numpy(m, tables_np)
    ptrs = R.table_pointers(tables)
    cases = {}
    for batch in (64, 257):
        req = pack_request(m, batch, synth.make_features(m, batch, seed=batch))
        ref, _ = run_oracle(m, req, tables_np)
        cases[batch] = (req, R.upload_blob(req), plan.alloc_out(batch),
                        plan.alloc_workspace(batch, req.sparse_ids), ref)
    first = {}
    for it in range(40):
        batch = (64, 257)[it % 2] if it % 5 else 64
        req, blob, out, ws, ref = cases[batch]
        out.fill_(float("nan"))
        plan.run(req, blob, ptrs, out, ws)
        torch.cuda.synchronize()
        res = out.float().cpu().numpy()
        assert not np.isnan(res).any()
        if batch not in first:
            first[batch] = res
            assert np.abs(res - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
        else:
            np.testing.assert_array_equal(res, first[batch])
    plan.close()


def test_argument_cache_is_keyed_by_content():
    """A request seen twice is served from a device-resident copy of its argument block; a
    request that differs only in its table pointers must not hit that copy."""
    import torch
    from recom_b200 import runtime as R
    m = synth.model_E_small()
    plan = R.Plan(m)
    batch = 80
    req = pack_request(m, batch, synth.make_features(m, batch, seed=1))
    blob = R.upload_blob(req)
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    outs = {}
    for tag, seed in (("a", 0), ("b", 7), ("a2", 0)):
        tables_np = synth.make_tables_np(m, seed=seed)
        tables = R.tables_from_numpy(m, tables_np)   # fresh device buffers: new pointers
        ptrs = R.table_pointers(tables)
        out = plan.alloc_out(batch)
        for _ in range(4):                           # second sight publishes, later ones hit
            out.fill_(float("nan"))
            plan.run(req, blob, ptrs, out, ws)
            torch.cuda.synchronize()
        ref, _ = run_oracle(m, req, tables_np)
        res = out.float().cpu().numpy()
        assert np.abs(res - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), tag
        outs[tag] = res
    assert np.abs(outs["a"] - outs["b"]).max() > 0
    np.testing.assert_array_equal(outs["a"], outs["a2"])
    plan.close()


@pytest.mark.parametrize("batch", [1, 31, 33, 2048])
def test_fast_runs_and_queue_tiles_cover_every_row(batch):
    """Row ranges are cut into runs (static) and queue tiles (dynamic); every output element
    must be written exactly once at any batch size, including partial tiles."""
    m = synth.dlrm_model(40, 9, 9, 3, 1, large_rows=1 << 14)
    feats = synth.make_features(m, batch, seed=batch, empty_frac=0.1)
    assert_parity(m, batch, feats, synth.make_tables_np(m))
