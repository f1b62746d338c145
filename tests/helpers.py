"""Shared helpers of the GPU parity tests: run a model through the C ABI on cuda:0 and
through the CPU oracle on the same packed request, and compare."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from recom_b200 import _lib as L
from recom_b200.schema import Model, PackedRequest, pack_request


def oracle_tables(model: Model, tables_np):
    """fp32 host tables -> what the oracle consumes (bf16 models: uint16 bit patterns)."""
    if model.out_dtype == L.BF16:
        return [O.f32_to_bf16_bits(t) for t in tables_np]
    return [np.ascontiguousarray(t, np.float32) for t in tables_np]


def run_oracle(model: Model, req: PackedRequest, tables_np, threads: int = 0):
    descs, bnd = model.descs()
    return O.run_model(descs, bnd, model.out_width, model.out_dtype, req.batch, req.blob,
                       req.slot_offsets, req.slot_elems, oracle_tables(model, tables_np),
                       req.total_ids, threads)


def run_gpu(model: Model, req: PackedRequest, tables_np, check_ids=False, want_ids=True,
            device=0, plan=None):
    import torch
    from recom_b200 import runtime as R

    own = plan is None
    plan = plan or R.Plan(model, device=device, check_ids=check_ids)
    tables = R.tables_from_numpy(model, tables_np, device)
    ptrs = R.table_pointers(tables)
    blob = R.upload_blob(req, device)
    out = plan.alloc_out(req.batch)
    out.fill_(float("nan"))
    ws = plan.alloc_workspace(req.batch, req.sparse_ids)
    ids = torch.full((max(req.total_ids, 1),), -7, dtype=torch.int32, device=f"cuda:{device}") if want_ids else None
    status = torch.zeros(1, dtype=torch.int32, device=f"cuda:{device}")
    plan.run(req, blob, ptrs, out, ws, ids_out=ids, status_out=status)
    torch.cuda.synchronize(device)
    launches = plan.last_launch_count()
    res = out.float().cpu().numpy()
    ids_np = ids.cpu().numpy()[: req.total_ids] if want_ids else None
    st = int(status.item())
    if own:
        plan.close()
    return res, ids_np, st, launches


def assert_parity(model: Model, batch: int, feats, tables_np, check_ids=False):
    req = pack_request(model, batch, feats)
    ref, ref_ids = run_oracle(model, req, tables_np)
    got, got_ids, status, launches = run_gpu(model, req, tables_np, check_ids=check_ids)
    assert status == 0
    assert launches >= (1 if batch > 0 else 0)
    # ids: bit exact
    np.testing.assert_array_equal(got_ids.astype(np.int64), ref_ids)
    assert not np.isnan(got).any(), "some output element was never written"
    # Tolerance (BASELINE.json): 1e-5 relative for fp32, 1e-2 for bf16. "Relative" is taken
    # against the magnitude of what was summed (the oracle re-run on |table|), the standard
    # forward bound for a sum whose order differs (TF-CPU adds 8-blocks, the GPU adds in
    # sequence); a plain |got-ref|/|ref| is meaningless where terms cancel.
    scale, _ = run_oracle(model, req, [np.abs(t) for t in tables_np])
    rtol = 1e-2 if model.out_dtype == L.BF16 else 1e-5
    err = np.abs(got - ref)
    bound = rtol * scale + 1e-30
    worst = float((err / bound).max()) if err.size else 0.0
    assert worst <= 1.0, f"max err/bound = {worst:.3g} (max abs err {err.max():.3g})"
    return req, got, ref
