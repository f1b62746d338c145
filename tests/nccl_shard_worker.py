"""torchrun worker of tests/test_gpu_sharding.py::test_nccl_all_gather_two_gpus."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from recom_b200 import runtime as R  # noqa: E402
from recom_b200 import sharding, synth  # noqa: E402
from recom_b200.schema import pack_request  # noqa: E402
from tests.helpers import run_oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    m = synth.mixed_model(64, seed=13, large_rows=1 << 14)
    batch = 96
    feats = synth.make_features(m, batch, seed=6, empty_frac=0.1)
    tables = synth.make_tables_np(m)
    shards = sharding.partition_columns(m, world)
    s = shards[rank]
    plan = R.Plan(s.model, device=local)
    dev_tables = R.tables_from_numpy(s.model, [tables[t] for t in s.table_ids], local)
    req = pack_request(s.model, batch, sharding.shard_features(s, feats))
    out = plan.alloc_out(batch)
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    plan.run(req, R.upload_blob(req, local), R.table_pointers(dev_tables), out, ws)
    full = sharding.all_gather_columns(out, shards)
    torch.cuda.synchronize()
    ref, _ = run_oracle(m, pack_request(m, batch, feats), tables)
    np.testing.assert_allclose(full.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    mine = sharding.all_to_all_columns(out, shards)
    torch.cuda.synchronize()
    lo, hi = sharding.batch_rows(batch, world, rank)
    np.testing.assert_array_equal(mine.cpu().numpy(), full.cpu().numpy()[lo:hi])
    # exchange fused into the kernel: rows stored straight into their owners over NVLink
    peers = sharding.PeerOutputs(shards, batch, m.out_width, 4)
    mine_t = peers.local_tensor(torch.float32)
    mine_t.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    # two steps: no host barrier or collective between launch and read, only the flag wait
    for _ in range(2):
        plan.run_sharded(req, R.upload_blob(req, local), R.table_pointers(dev_tables), ws, peers.next_step())
        peers.wait()
        got = mine_t.clone()
    torch.cuda.synchronize()
    assert not peers.timed_out()
    # the barrier inside the kernel: no wait at all behind the launch
    dist.barrier()
    mine_t.fill_(float("nan"))
    torch.cuda.synchronize()
    dist.barrier()
    plan.run_sharded(req, R.upload_blob(req, local), R.table_pointers(dev_tables), ws, peers.next_step(True))
    got2 = mine_t.clone()
    torch.cuda.synchronize()
    assert not peers.timed_out()
    assert torch.equal(got2[: got.shape[0]].nan_to_num(), got.nan_to_num())
    rpr = peers.rows_per_rank
    lo, hi = rank * rpr, min((rank + 1) * rpr, batch)
    np.testing.assert_array_equal(got.cpu().numpy()[: hi - lo], full.cpu().numpy()[lo:hi])
    peers.close()
    dist.barrier()
    # Mode B: table rows sharded over the ranks, partial sums reduce-scattered (NCCL), scaled after
    stage = sharding.RowShardedStage(m, world, rank, local)
    tabs_b = stage.local_tables(tables)
    ptrs_b = R.table_pointers(tabs_b)
    req_b = pack_request(m, batch, feats)
    blob_b = R.upload_blob(req_b, local)
    out_b = stage.plan.alloc_out(batch)
    ws_b = stage.plan.alloc_workspace(batch, req_b.sparse_ids)
    stage.partial(req_b, blob_b, ptrs_b, out_b, ws_b)
    mine_b, row0 = stage.finish(req_b, blob_b, ptrs_b, ws_b, out_b)
    torch.cuda.synchronize()
    want = ref[row0:row0 + mine_b.shape[0]]
    err = np.abs(mine_b.cpu().numpy() - want)
    assert np.all(err <= 2e-5 * np.maximum(np.abs(want), 1.0)), float(err.max())
    dist.barrier()
    if rank == 0:
        print("SHARD_OK MODE_B_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
