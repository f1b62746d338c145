"""Multi-GPU host logic on CPU: the column-group partitioner, the row-shard decomposition and
the exchange step (world_size 2 over gloo). Shard results are computed with the CPU oracle
(test infrastructure) and must reassemble to the single-device result."""
import os
import socket

import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import sharding, synth
from recom_b200.schema import pack_request
from tests.helpers import run_oracle


def _full(model, batch, feats, tables):
    return run_oracle(model, pack_request(model, batch, feats), tables)[0]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_is_contiguous_and_complete(world):
    m = synth.mixed_model(101, seed=5, large_rows=1 << 12)
    shards = sharding.partition_columns(m, world)
    assert len(shards) == world
    assert shards[0].first_col == 0 and shards[-1].last_col == len(m.columns)
    for a, b in zip(shards, shards[1:]):
        assert a.last_col == b.first_col and a.out_end == b.out_begin
    assert shards[-1].out_end == m.out_width
    assert sum(s.width for s in shards) == m.out_width
    es = 4
    costs = [sum(sharding.column_cost(c, es) for c in m.columns[s.first_col:s.last_col]) for s in shards]
    assert max(costs) <= 1.6 * (sum(costs) / world) + max(sharding.column_cost(c, es) for c in m.columns)
    # local models are self-contained: local table indices, offsets from zero
    for s in shards:
        assert s.model.out_width == s.width
        assert len(s.table_ids) == len(s.model.table_shapes)
        for c, g in zip(s.model.columns, m.columns[s.first_col:s.last_col]):
            assert s.table_ids[c.table] == g.table and c.out_offset == g.out_offset - s.out_begin


def test_partition_respects_table_budget():
    m = synth.large_table_model(16, rows=1000, dim=32)
    shards = sharding.partition_columns(m, 4, max_table_bytes=4 * 1000 * 32 * 4)
    assert all(len(s.model.columns) == 4 for s in shards)
    with pytest.raises(AssertionError):
        sharding.partition_columns(m, 2, max_table_bytes=4 * 1000 * 32 * 4)


@pytest.mark.parametrize("world", [2, 4])
def test_column_shards_reassemble_to_single_device_result(world):
    m = synth.mixed_model(60, seed=7, large_rows=1 << 12)
    batch = 37
    feats = synth.make_features(m, batch, seed=1, empty_frac=0.1)
    tables = synth.make_tables_np(m)
    ref = _full(m, batch, feats, tables)
    outs = []
    shards = sharding.partition_columns(m, world)
    for s in shards:
        lt = [tables[t] for t in s.table_ids]
        outs.append(_full(s.model, batch, sharding.shard_features(s, feats), lt))
    np.testing.assert_array_equal(sharding.assemble_host(shards, outs), ref)  # Mode A: bit identical


@pytest.mark.parametrize("combiner", [L.COMBINE_SUM, L.COMBINE_MEAN, L.COMBINE_SQRTN])
def test_row_shards_sum_to_single_device_result(combiner):
    from recom_b200.schema import Column, Feature, Model
    rng = np.random.default_rng(3)
    rows, dim, batch, world = 1000, 16, 50, 4
    col = Column("c", L.LOOKUP_SEGMENT, L.XFORM_NONE, L.I64, 0, rows, dim, combiner=combiner,
                 seg_format=L.SEG_ROW_SPLITS_I32)
    m = Model([col], [(rows, dim)]).finalize()
    table = rng.standard_normal((rows, dim)).astype(np.float32)
    lengths = synth.row_lengths(rng, batch, "u10", empty_frac=0.2)
    rs, coo = synth.splits_and_coo(lengths)
    ids = rng.integers(0, rows, size=int(rs[-1])).astype(np.int64)
    ref = _full(m, batch, {"c": Feature(ids, row_splits=rs, coo=coo)}, [table])
    row_of = coo[:, 0]
    partial = np.zeros((batch, dim), np.float64)
    total = None
    for r in range(world):
        sh = sharding.RowShard(r, world)
        lm = sh.local_model(m)
        feat, total = sh.local_feature(ids, row_of, batch)
        out = _full(lm, batch, {lm.columns[0].key: feat}, [sh.local_table(table)])
        partial += out
    got = partial * sharding.combine_scale(combiner, total)[:, None]
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)  # Mode B: reduction-order tolerance


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = synth.mixed_model(40, seed=9, large_rows=1 << 12)
        batch = 24
        feats = synth.make_features(m, batch, seed=2, empty_frac=0.1)  # replicated inputs
        tables = synth.make_tables_np(m)
        shards = sharding.partition_columns(m, world)
        s = shards[rank]
        local = _full(s.model, batch, sharding.shard_features(s, feats), [tables[t] for t in s.table_ids])
        full = sharding.all_gather_columns(torch.from_numpy(local), shards).numpy()
        ref = _full(m, batch, feats, tables)
        mine = sharding.all_to_all_columns(torch.from_numpy(local), shards).numpy()
        lo, hi = sharding.batch_rows(batch, world, rank)
        ok = bool(np.array_equal(full, ref)) and bool(np.array_equal(mine, ref[lo:hi]))
        q.put((rank, ok, full.shape))
    finally:
        dist.destroy_process_group()


def test_all_gather_columns_world2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(ok for _, ok, _ in results)
