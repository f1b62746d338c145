"""GPU parity at the sizes BASELINE.json quotes its numbers on: the CUDA path (through the C
ABI) against the CPU oracle on the same packed request, full batch, every column.

  * ids bit-exact (all columns);
  * DENSE / GATHER_SCATTER slices are pure copies of table rows: bit-exact;
  * combined embeddings within north_star's tolerance - both the forward bound
    |err| <= rtol * sum|terms| (tests/helpers.py) and plain numpy allclose
    (rtol 1e-5, atol 1e-6 for fp32; rtol 1e-2, atol 1e-5 for bf16).

Tables are generated on the host (torch CPU generator, fixed seed) and uploaded, so the oracle
and the device read the same bits. The oracle does model E at B=2048 in a few ms per pass; the
cost of these tests is table generation and upload (a few seconds per model)."""
import numpy as np
import pytest

from recom_b200 import _lib as L
from recom_b200 import synth
from recom_b200.schema import pack_request
from tests.helpers import run_gpu, run_oracle

pytestmark = pytest.mark.gpu


def host_tables(model, seed=0):
    import torch

    g = torch.Generator()
    g.manual_seed(seed + 777)
    out = []
    for rows, dim in model.table_shapes:
        t = torch.empty((rows, dim), dtype=torch.float32)
        t.normal_(0.0, 1.0 / float(np.sqrt(dim)), generator=g)
        out.append(t.numpy())
    return out


class Rig:
    """One model resident on the device and on the host; runs any number of batches."""

    def __init__(self, model):
        from recom_b200 import runtime as R

        self.model = model
        self.tables_np = host_tables(model)
        self.plan = R.Plan(model)
        self.tables = R.tables_from_numpy(model, self.tables_np)
        self.ptrs = R.table_pointers(self.tables)
        if model.out_dtype == L.BF16:
            # the oracle reads bf16 bit patterns; the device tables were rounded the same way
            from oracle import oracle as O
            self.oracle_tables = [O.f32_to_bf16_bits(t) for t in self.tables_np]
            self.abs_tables = [O.f32_to_bf16_bits(np.abs(t)) for t in self.tables_np]
        else:
            self.oracle_tables = self.tables_np
            self.abs_tables = None  # made lazily, in place, after the last reference run

    def oracle(self, req, tables):
        from oracle import oracle as O

        m = self.model
        descs, bnd = m.descs()
        return O.run_model(descs, bnd, m.out_width, m.out_dtype, req.batch, req.blob,
                           req.slot_offsets, req.slot_elems, tables, req.total_ids, 0)

    def device(self, req):
        import torch
        from recom_b200 import runtime as R

        out = self.plan.alloc_out(req.batch)
        out.fill_(float("nan"))
        ws = self.plan.alloc_workspace(req.batch, req.sparse_ids)
        ids = torch.full((max(req.total_ids, 1),), -7, dtype=torch.int32, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        blob = R.upload_blob(req)
        self.plan.run(req, blob, self.ptrs, out, ws, ids_out=ids, status_out=status)
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        # a second run without the parity hook must give the same bits (the timed path)
        out2 = self.plan.alloc_out(req.batch)
        out2.fill_(float("nan"))
        self.plan.run(req, blob, self.ptrs, out2, ws)
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.uint8), out2.view(torch.uint8))
        return out.float().cpu().numpy(), ids.cpu().numpy()[: req.total_ids]

    def check(self, batches, seed0=0, plain=True, **feat_kw):
        m = self.model
        bf16 = m.out_dtype == L.BF16
        rtol = 1e-2 if bf16 else 1e-5
        cases = []
        for k, B in enumerate(batches):
            req = pack_request(m, B, synth.make_features(m, B, seed=seed0 + k, **feat_kw))
            ref, ref_ids = self.oracle(req, self.oracle_tables)
            got, got_ids = self.device(req)
            np.testing.assert_array_equal(got_ids.astype(np.int64), ref_ids)
            assert not np.isnan(got).any(), "some output element was never written"
            for c in m.columns:
                sl = slice(c.out_offset, c.out_offset + c.dim)
                if c.lookup != L.LOOKUP_SEGMENT:
                    np.testing.assert_array_equal(got[:, sl], ref[:, sl], err_msg=c.name)
            atol = 1e-5 if bf16 else 1e-6
            bad = ~np.isclose(got, ref, rtol=rtol, atol=atol)
            # (not for rows of thousands of ids: the rounding error of a long fp32 sum taken in
            # another order is bounded by the summed magnitudes, not by the result)
            assert not (plain and bad.any()), f"B={B}: {bad.sum()} elements outside rtol={rtol} atol={atol}"
            cases.append((req, got, ref))
        # forward bound against the summed magnitudes (the oracle re-run on |table|)
        if self.abs_tables is None:
            for t in self.tables_np:
                np.abs(t, out=t)
            self.abs_tables = self.tables_np
        for req, got, ref in cases:
            scale, _ = self.oracle(req, self.abs_tables)
            worst = float((np.abs(got - ref) / (rtol * scale + 1e-30)).max())
            assert worst <= 1.0, f"B={req.batch}: max err/bound = {worst:.3g}"

    def close(self):
        self.plan.close()


def test_model_E_b2048_b4096_b8192_vs_oracle():
    rig = Rig(synth.model_E())
    rig.check([2048, 4096, 8192, 512, 1])
    rig.close()


def test_model_F_b2048_vs_oracle():
    rig = Rig(synth.model_F())
    rig.check([2048, 4096])
    rig.close()


@pytest.mark.parametrize("out_dtype", [L.F32, L.BF16])
def test_mixed1000_b2048_vs_oracle(out_dtype):
    rig = Rig(synth.mixed_model(1000, seed=0, out_dtype=out_dtype))
    rig.check([2048], seed0=3)
    rig.close()


def test_production2000_b2048_vs_oracle():
    rig = Rig(synth.production_model(2000, share=4, seed=0))
    rig.check([2048], seed0=5)
    rig.close()


def test_model_E_bf16_b2048_vs_oracle():
    rig = Rig(synth.model_E(out_dtype=L.BF16, large_rows=1 << 21))
    rig.check([2048], seed0=7)
    rig.close()


@pytest.mark.parametrize("dist", ["u50", "zipf"])
def test_dynamic_shape_b8192_vs_oracle(dist):
    """BASELINE config 3: per-request segment counts, long tails and empty rows at the top of
    the batch sweep."""
    rig = Rig(synth.dlrm_model(40, 5, 5, 15, 5, large_rows=1 << 20, name="dyn"))
    rig.check([8192, 33], seed0=11, plain=dist != "zipf", dist=dist, empty_frac=0.2)
    rig.close()


def test_scatter_columns_host_inputs():
    """recom_scatter_columns (ConcatOutputs' host_inputs capability,
    concat_outputs_op_gpu.cu.cc:186-232): columns computed elsewhere land in their slices of
    the concat tensor and nothing else is touched."""
    import ctypes as C

    import torch

    rng = np.random.default_rng(0)
    for dt, eb in ((torch.float32, 4), (torch.bfloat16, 2)):
        for batch in (1, 77, 2048):
            width = 53
            dims = [3, 8, 1, 16]
            offs = [0, 5, 20, 37]
            out = torch.full((batch, width), -1.0, dtype=dt, device="cuda")
            cols = [torch.from_numpy(rng.standard_normal((batch, d)).astype(np.float32)).cuda().to(dt)
                    for d in dims]
            ptrs = np.array([c.data_ptr() for c in cols], dtype=np.uint64)
            d_arr = np.array(dims, np.int32)
            o_arr = np.array(offs, np.int32)
            L.check(L.lib.recom_scatter_columns(ptrs.ctypes.data, d_arr.ctypes.data, o_arr.ctypes.data,
                                                len(dims), batch, width, eb, out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            want = torch.full((batch, width), -1.0, dtype=dt, device="cuda")
            for c, d, o in zip(cols, dims, offs):
                want[:, o:o + d] = c
            assert torch.equal(out, want)
    # argument errors are reported, not fatal
    bad = np.array([60], np.int32)
    one = np.array([4], np.int32)
    p = np.array([cols[0].data_ptr()], dtype=np.uint64)
    st = L.lib.recom_scatter_columns(p.ctypes.data, one.ctypes.data, bad.ctypes.data, 1, 4, 53, 4,
                                     out.data_ptr(), None)
    assert st == L.ERR_INVALID_ARGUMENT
