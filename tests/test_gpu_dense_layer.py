"""the first dense layer behind the concat (recom_dense_relu, a tcgen05 / TMA
kernel) against a plain PyTorch fp32 reference of the same op on the same bf16 operands: the
products of bf16 values are exact in fp32, so the two differ only by the order of the fp32 sum -
tolerance 2e-5 of sum |a_k w_k| (+ 1e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference(a, w, bias, relu):
    import torch

    ref = a.double() @ w.double().T
    if bias is not None:
        ref = ref + bias.double()
    if relu:
        ref = ref.clamp_min(0)
    scale = a.double().abs() @ w.double().abs().T
    return ref, scale


@pytest.mark.parametrize("m,k,n", [(128, 64, 128), (128, 256, 128), (2048, 8120, 1024), (300, 520, 200),
                                   (1, 8, 1), (129, 72, 130), (257, 4096, 16),
                                   # 128 x 256 tiles (grids that still fill the GPU), full and ragged in N
                                   (4096, 520, 1024), (5000, 264, 600)])
@pytest.mark.parametrize("relu", [True, False])
def test_dense_relu_matches_fp32_reference(m, k, n, relu):
    import torch
    from recom_b200 import runtime as R

    g = torch.Generator(device="cuda").manual_seed(m * 131 + k)
    a = torch.randn((m, k), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((n, k), device="cuda", generator=g) / np.sqrt(k)).to(torch.bfloat16)
    bias = torch.randn((n,), device="cuda", generator=g)
    out = torch.full((m, n), float("nan"), device="cuda")
    R.dense_relu(a, w, bias, relu=relu, out=out)
    torch.cuda.synchronize()
    ref, scale = _reference(a, w, bias, relu)
    err = (out.double() - ref).abs()
    assert not torch.isnan(out).any(), "some output element was never written"
    bound = 2e-5 * scale + 1e-6
    assert bool((err <= bound).all()), f"max err {err.max().item():.3g}, worst err/bound {(err / bound).max().item():.3g}"


def test_dense_layer_behind_a_bf16_plan():
    """The real consumer: the bf16 concat rows a plan has just written (row stride = out_width)."""
    import torch
    from recom_b200 import consts as L
    from recom_b200 import runtime as R
    from recom_b200 import synth
    from recom_b200.schema import pack_request

    model = synth.mixed_model(96, seed=4, out_dtype=L.BF16, large_rows=1 << 14)
    batch = 500
    feats = synth.make_features(model, batch, seed=2)
    req = pack_request(model, batch, feats)
    plan = R.Plan(model)
    tables = R.tables_from_numpy(model, synth.make_tables_np(model))
    blob = R.upload_blob(req)
    out = plan.alloc_out(batch)
    ws = plan.alloc_workspace(batch, req.sparse_ids)
    plan.run(req, blob, R.table_pointers(tables), out, ws)
    k = model.out_width
    assert k % 8 == 0 and out.dtype == torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(7)
    w = (torch.randn((512, k), device="cuda", generator=g) / np.sqrt(k)).to(torch.bfloat16)
    y = R.dense_relu(out, w, None, relu=True)
    torch.cuda.synchronize()
    ref, scale = _reference(out, w, None, True)
    assert bool(((y.double() - ref).abs() <= 2e-5 * scale + 1e-6).all())
    plan.close()


def test_dense_relu_rejects_bad_operands():
    import torch
    from recom_b200 import _lib as L
    from recom_b200 import runtime as R

    a = torch.zeros((4, 12), device="cuda", dtype=torch.bfloat16)  # K not a multiple of 8
    w = torch.zeros((4, 12), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(Exception):
        R.dense_relu(a, w)
