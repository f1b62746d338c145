"""GPU: shard-local plans reproduce the single-device result (Mode A bit-identical), and the
NCCL exchange when two GPUs are visible."""
import numpy as np
import pytest

from recom_b200 import sharding, synth
from recom_b200.schema import pack_request
from tests.helpers import run_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4])
def test_column_shards_on_one_gpu_reassemble(world):
    m = synth.mixed_model(80, seed=11, large_rows=1 << 14)
    batch = 130
    feats = synth.make_features(m, batch, seed=4, empty_frac=0.1)
    tables = synth.make_tables_np(m)
    full, ids, status, _ = run_gpu(m, pack_request(m, batch, feats), tables)
    assert status == 0
    outs = []
    shards = sharding.partition_columns(m, world)
    for s in shards:
        req = pack_request(s.model, batch, sharding.shard_features(s, feats))
        o, _, st, _ = run_gpu(s.model, req, [tables[t] for t in s.table_ids])
        assert st == 0
        outs.append(o)
    np.testing.assert_array_equal(sharding.assemble_host(shards, outs), full)


@pytest.mark.parametrize("world,batch,bf16", [(2, 130, False), (4, 257, False), (8, 256, False), (3, 100, True)])
def test_fused_exchange_addressing_on_one_gpu(world, batch, bf16):
    """recom_plan_run_sharded with every rank's concat tensor on this device: each shard's
    kernel stores its column slice of every row into the row's owner; the owners' tensors,
    stacked, are bit-identical to the single-plan result."""
    import ctypes as C

    import torch
    from recom_b200 import _lib as L
    from recom_b200 import runtime as R

    m = synth.mixed_model(96, seed=17, large_rows=1 << 14, out_dtype=L.BF16) if bf16 else \
        synth.mixed_model(96, seed=17, large_rows=1 << 14)
    feats = synth.make_features(m, batch, seed=5, empty_frac=0.1)
    tables = synth.make_tables_np(m)
    full, _, status, _ = run_gpu(m, pack_request(m, batch, feats), tables)
    assert status == 0
    shards = sharding.partition_columns(m, world)
    rpr = max(32, (batch + world - 1) // world)
    dt = torch.bfloat16 if bf16 else torch.float32
    owners = [torch.full((rpr, m.out_width), float("nan"), dtype=dt, device="cuda:0") for _ in range(world)]
    ptrs = np.array([o.data_ptr() for o in owners], dtype=np.uint64)
    flags = [torch.zeros(L.SHARD_FLAG_WORDS, dtype=torch.int32, device="cuda:0") for _ in range(world)]
    fptrs = np.array([f.data_ptr() for f in flags], dtype=np.uint64)
    for s in shards:
        plan = R.Plan(s.model, device=0)
        dev_tables = R.tables_from_numpy(s.model, [tables[t] for t in s.table_ids], 0)
        req = pack_request(s.model, batch, sharding.shard_features(s, feats))
        ws = plan.alloc_workspace(batch, req.sparse_ids)
        so = L.ShardOut(world=world, col_offset=s.out_begin, global_width=m.out_width, rank=s.rank,
                        rows_per_rank=rpr, peer_out=ptrs.ctypes.data, peer_flags=fptrs.ctypes.data,
                        epoch=7, flags=0)
        plan.run_sharded(req, R.upload_blob(req, 0), R.table_pointers(dev_tables), ws, so)
        torch.cuda.synchronize()
        plan.close()
    # every rank's kernel signalled every owner; the stream-side wait returns at once
    for f in flags:
        L.check(L.lib.recom_shard_wait(f.data_ptr(), world, 7, 2000, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    for f in flags:
        v = f.cpu().numpy()
        assert (v[:world] == 7).all() and (v[world:] == 0).all()
    got = torch.cat(owners, dim=0)[:batch].float().cpu().numpy()
    np.testing.assert_array_equal(got, full)
    # a signal that never comes: the wait gives up and says so
    L.check(L.lib.recom_shard_wait(flags[0].data_ptr(), world, 8, 20, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert int(flags[0][L.SHARD_FLAG_WORDS - 1].item()) == 1


def test_fused_exchange_world1_waits_in_kernel():
    """World 1 with RECOM_SHARD_WAIT_IN_KERNEL: the launch signals itself, sees its own flag and
    ends (the multi-rank case needs concurrent kernels: tests/nccl_shard_worker.py)."""
    import torch
    from recom_b200 import _lib as L
    from recom_b200 import runtime as R

    m = synth.mixed_model(40, seed=5, large_rows=1 << 12)
    batch = 77
    feats = synth.make_features(m, batch, seed=2, empty_frac=0.2)
    tables = synth.make_tables_np(m)
    req = pack_request(m, batch, feats)
    full, _, status, _ = run_gpu(m, req, tables)
    plan = R.Plan(m, device=0)
    dev_tables = R.tables_from_numpy(m, tables, 0)
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    out = torch.full((80, m.out_width), float("nan"), device="cuda:0")
    flags = torch.zeros(L.SHARD_FLAG_WORDS, dtype=torch.int32, device="cuda:0")
    ptrs = np.array([out.data_ptr()], dtype=np.uint64)
    fptrs = np.array([flags.data_ptr()], dtype=np.uint64)
    blob = R.upload_blob(req, 0)
    for epoch in (1, 2, 3):
        so = L.ShardOut(world=1, col_offset=0, global_width=m.out_width, rank=0, rows_per_rank=80,
                        peer_out=ptrs.ctypes.data, peer_flags=fptrs.ctypes.data, epoch=epoch,
                        flags=L.SHARD_WAIT_IN_KERNEL)
        plan.run_sharded(req, blob, R.table_pointers(dev_tables), ws, so)
    torch.cuda.synchronize()
    v = flags.cpu().numpy()
    assert v[0] == 3 and v[L.SHARD_FLAG_WORDS - 1] == 0
    np.testing.assert_array_equal(out[:batch].cpu().numpy(), full)
    plan.close()


def test_fused_exchange_rejects_bad_arguments():
    import torch
    from recom_b200 import _lib as L
    from recom_b200 import runtime as R

    m = synth.mixed_model(16, seed=3, large_rows=1 << 12)
    batch = 64
    feats = synth.make_features(m, batch, seed=1)
    req = pack_request(m, batch, feats)
    plan = R.Plan(m, device=0)
    dev_tables = R.tables_from_numpy(m, synth.make_tables_np(m), 0)
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    out = torch.zeros((64, m.out_width), device="cuda:0")
    ptrs = np.array([out.data_ptr()], dtype=np.uint64)
    blob = R.upload_blob(req, 0)
    for kw in (dict(world=0), dict(world=9), dict(rank=1), dict(rank=-1), dict(rows_per_rank=16), dict(rows_per_rank=32),
               dict(col_offset=4), dict(global_width=m.out_width - 4)):
        a = dict(world=1, col_offset=0, global_width=m.out_width, rank=0, rows_per_rank=64,
                 peer_out=ptrs.ctypes.data, peer_flags=None, epoch=0, flags=0)
        a.update(kw)
        with pytest.raises(L.RecomError):
            plan.run_sharded(req, blob, R.table_pointers(dev_tables), ws, L.ShardOut(**a))
    plan.close()


def test_nccl_all_gather_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(root, "tests", "nccl_shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARD_OK" in out.stdout


def test_mode_b_row_sharded_tables_single_gpu():
    """Mode B on the device: the table ROWS of every column sharded over 3 ranks,
    all ranks emulated on cuda:0 - every rank's plan skips the ids of other ranks' rows inside
    the kernel and writes partial sums; their sum, scaled by recom_plan_scale_rows, equals the
    single-plan result (<= 1e-5 relative) for hashed, bucketized, multi-hot mean / sqrtn and
    one-id columns alike."""
    import torch

    from recom_b200 import runtime as R
    from recom_b200.sharding import RowShardedStage

    model = synth.mixed_model(60, seed=3, large_rows=1 << 12)
    batch = 96
    feats = synth.make_features(model, batch, seed=5, empty_frac=0.1)
    tables_np = synth.make_tables_np(model)
    req = pack_request(model, batch, feats)
    # single-plan reference
    full, _, status, _ = run_gpu(model, req, tables_np, want_ids=False)
    assert status == 0
    world = 3
    blob = R.upload_blob(req, 0)
    total = None
    stages = []
    for g in range(world):
        st = RowShardedStage(model, world, g, 0)
        tabs = st.local_tables(tables_np)
        ptrs = R.table_pointers(tabs)
        out = st.plan.alloc_out(batch)
        ws = st.plan.alloc_workspace(batch, req.sparse_ids)
        st.partial(req, blob, ptrs, out, ws)
        torch.cuda.synchronize()
        total = out.clone() if total is None else total + out
        stages.append((st, ptrs, ws, tabs))
    st, ptrs, ws, _ = stages[0]
    st.plan.scale_rows(req, blob, ptrs, ws, total, 0)
    torch.cuda.synchronize()
    got = total.float().cpu().numpy()
    scale = np.abs(full) + 1e-6
    assert np.all(np.abs(got - full) <= 2e-5 * np.maximum(scale, 1.0)), float(np.abs(got - full).max())
