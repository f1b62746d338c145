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
