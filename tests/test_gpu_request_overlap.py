"""Back-to-back requests under programmatic dependent launch (recom_plan_set_request_overlap): the
prologue of request i + 1 overlaps the tail of request i, nothing of i + 1 is stored before i is
complete. Many requests in a row - into rotating output buffers and into ONE reused buffer - must
give exactly what the fully ordered launches give."""
import numpy as np
import pytest

from recom_b200 import synth
from recom_b200.schema import pack_request

pytestmark = pytest.mark.gpu


def test_overlapped_requests_equal_ordered_requests():
    import torch
    from recom_b200 import runtime as R

    model = synth.dlrm_model(880, 10, 10, 4, 0, name="E_like")  # a grid that fills the GPU, one fused launch
    batch, n_req = 1024, 6
    plan = R.Plan(model)
    tables_np = synth.make_tables_np(model)
    tables = R.tables_from_numpy(model, tables_np)
    ptrs = R.table_pointers(tables)
    reqs = [pack_request(model, batch, synth.make_features(model, batch, seed=10 + i)) for i in range(n_req)]
    blobs = [R.upload_blob(r) for r in reqs]
    wss = [plan.alloc_workspace(batch, r.sparse_ids) for r in reqs]

    def run_all(outs, rounds):
        for _ in range(rounds):
            for i in range(n_req):
                plan.run(reqs[i], blobs[i], ptrs, outs[i], wss[i])
        torch.cuda.synchronize()
        return [o.clone() for o in outs]

    plan.set_request_overlap(False)
    ref = run_all([plan.alloc_out(batch) for _ in range(n_req)], 1)
    assert plan.last_launch_count() == 1
    plan.set_request_overlap(True)
    got = run_all([plan.alloc_out(batch).fill_(float("nan")) for _ in range(n_req)], 5)
    for g, r in zip(got, ref):
        assert torch.equal(g, r)
    # one reused output buffer: the last request's result must be what is left in it
    one = plan.alloc_out(batch)
    for _ in range(5):
        for i in range(n_req):
            plan.run(reqs[i], blobs[i], ptrs, one, wss[i])
    torch.cuda.synchronize()
    assert torch.equal(one, ref[-1])
    plan.close()
