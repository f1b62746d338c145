"""Plan / request plumbing over the C ABI. torch is used only for device memory,
streams and events; every computation happens inside librecom_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .schema import Model, PackedRequest


def _torch_out_dtype(out_dtype: int) -> torch.dtype:
    return torch.bfloat16 if out_dtype == L.BF16 else torch.float32


class Plan:
    """Owns a recom_plan (the counterpart of the reference's generated .so + const buffer,
    feature_column_process_op_gpu.cu.cc:51-62)."""

    def __init__(self, model: Model, device: int = 0, check_ids: bool = False,
                 max_ctas_per_sm: int = 0, row_shard: Optional[tuple] = None):
        """row_shard = (world, rank): Mode B - the tables passed to run() are this rank's row shards
        (row r on rank r % world at local row r // world), the plan writes partial sums, and
        scale_rows() finishes the rows this rank keeps after the cross-rank reduction."""
        self.model = model
        self.device = device
        descs, boundaries = model.descs()
        self._descs, self._boundaries = descs, boundaries
        cfg = L.PlanConfig(out_dtype=model.out_dtype, out_width=model.out_width,
                           num_slots=model.num_slots, num_tables=len(model.table_shapes),
                           check_ids=int(check_ids), device=device,
                           max_ctas_per_sm=max_ctas_per_sm,
                           row_shard=(row_shard[0] | (row_shard[1] << 16)) if row_shard else 0)
        h = C.c_void_p()
        flt = model.filters()
        self._filters = flt
        L.check(L.lib.recom_plan_create_filtered(descs.ctypes.data, len(descs),
                                                 boundaries.ctypes.data if boundaries.size else None,
                                                 boundaries.size, flt.ctypes.data if flt is not None else None,
                                                 C.byref(cfg), C.byref(h)))
        self._h = h

    @classmethod
    def from_blob(cls, blob: bytes, model: Model, device: int = 0) -> "Plan":
        self = cls.__new__(cls)
        self.model, self.device = model, device
        h = C.c_void_p()
        L.check(L.lib.recom_plan_create_from_blob(blob, len(blob), device, C.byref(h)))
        self._h = h
        return self

    def serialize(self) -> bytes:
        need = C.c_size_t()
        L.check(L.lib.recom_plan_serialize(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        L.check(L.lib.recom_plan_serialize(self._h, buf, need.value, C.byref(need)))
        return buf.raw

    def close(self) -> None:
        if getattr(self, "_h", None):
            L.lib.recom_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace_bytes(self, batch: int, sparse_ids: int) -> int:
        out = C.c_size_t()
        L.check(L.lib.recom_plan_workspace_bytes(self._h, batch, sparse_ids, C.byref(out)))
        return out.value

    def _request(self, req: PackedRequest, blob_dev: torch.Tensor, tables_ptrs: np.ndarray,
                 out: torch.Tensor, workspace: torch.Tensor, ids_out, status_out, stream) -> L.Request:
        r = L.Request()
        r.batch = req.batch
        r.blob = blob_dev.data_ptr()
        r.slot_offsets = req.slot_offsets.ctypes.data
        r.slot_elems = req.slot_elems.ctypes.data
        r.tables = tables_ptrs.ctypes.data
        r.out = out.data_ptr()
        r.workspace = workspace.data_ptr()
        r.workspace_bytes = workspace.numel() * workspace.element_size()
        r.ids_out = ids_out.data_ptr() if ids_out is not None else None
        r.status_out = status_out.data_ptr() if status_out is not None else None
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        r.stream = s.cuda_stream
        r.args_offset = req.args_offset
        return r

    def run(self, req: PackedRequest, blob_dev: torch.Tensor, tables_ptrs: np.ndarray,
            out: torch.Tensor, workspace: torch.Tensor, ids_out: Optional[torch.Tensor] = None,
            status_out: Optional[torch.Tensor] = None, stream=None) -> None:
        """Enqueue the fused stage (stream ordered, returns immediately)."""
        r = self._request(req, blob_dev, tables_ptrs, out, workspace, ids_out, status_out, stream)
        L.check(L.lib.recom_plan_run(self._h, C.byref(r)))

    def run_sharded(self, req: PackedRequest, blob_dev: torch.Tensor, tables_ptrs: np.ndarray,
                    workspace: torch.Tensor, shard_out: "L.ShardOut", stream=None) -> None:
        """Mode A with the exchange fused into the kernel: rows are stored straight into the
        owning ranks' concat tensors (see sharding.PeerOutputs)."""
        r = self._request(req, blob_dev, tables_ptrs, workspace, workspace, None, None, stream)
        r.out = None
        L.check(L.lib.recom_plan_run_sharded(self._h, C.byref(r), C.byref(shard_out)))

    def scale_rows(self, req: PackedRequest, blob_dev: torch.Tensor, tables_ptrs: np.ndarray,
                   workspace: torch.Tensor, rows_tensor: torch.Tensor, row_begin: int, stream=None) -> None:
        """Mode B: mean / sqrtn scaling of rows [row_begin, row_begin + len(rows_tensor)) of the
        cross-rank sum (rows_tensor: [rows, out_width], contiguous), in place."""
        r = self._request(req, blob_dev, tables_ptrs, rows_tensor, workspace, None, None, stream)
        L.check(L.lib.recom_plan_scale_rows(self._h, C.byref(r), int(row_begin), int(rows_tensor.shape[0]),
                                            rows_tensor.data_ptr()))

    def algorithmic_bytes(self, req: PackedRequest, blob_dev, tables_ptrs, out, workspace) -> int:
        r = self._request(req, blob_dev, tables_ptrs, out, workspace, None, None, None)
        n = C.c_int64()
        L.check(L.lib.recom_plan_algorithmic_bytes(self._h, C.byref(r), C.byref(n)))
        return n.value

    def set_request_overlap(self, on: bool) -> None:
        """Programmatic dependent launch between consecutive requests (recom_plan_set_request_overlap)."""
        L.check(L.lib.recom_plan_set_request_overlap(self._h, 1 if on else 0))

    @staticmethod
    def last_launch_count() -> int:
        return L.lib.recom_last_launch_count()

    # ---- conveniences -------------------------------------------------------------
    def alloc_out(self, batch: int) -> torch.Tensor:
        return torch.empty((batch, self.model.out_width), dtype=_torch_out_dtype(self.model.out_dtype),
                           device=f"cuda:{self.device}")

    def alloc_workspace(self, batch: int, sparse_ids: int) -> torch.Tensor:
        n = self.workspace_bytes(batch, sparse_ids)
        return torch.empty(n + 256, dtype=torch.uint8, device=f"cuda:{self.device}")


def table_pointers(tables: Sequence[torch.Tensor]) -> np.ndarray:
    return np.array([t.data_ptr() for t in tables], dtype=np.uint64)


def make_tables(model: Model, device: int = 0, seed: int = 0) -> List[torch.Tensor]:
    """Random-init tables on the device: N(0, 1/sqrt(D)), the TF embedding_column default
    (truncated_normal(stddev=1/sqrt(dim))), in the model's table dtype."""
    g = torch.Generator(device=f"cuda:{device}")
    g.manual_seed(seed + 12345)
    dt = _torch_out_dtype(model.out_dtype)
    out = []
    for rows, dim in model.table_shapes:
        t = torch.empty((rows, dim), dtype=torch.float32, device=f"cuda:{device}")
        t.normal_(0.0, 1.0 / float(np.sqrt(dim)), generator=g)
        out.append(t.to(dt) if dt != torch.float32 else t)
    return out


def tables_from_numpy(model: Model, tables_np: Sequence[np.ndarray], device: int = 0) -> List[torch.Tensor]:
    dt = _torch_out_dtype(model.out_dtype)
    return [torch.from_numpy(np.ascontiguousarray(t)).to(f"cuda:{device}").to(dt).contiguous()
            for t in tables_np]


def upload_blob(req: PackedRequest, device: int = 0) -> torch.Tensor:
    return torch.from_numpy(req.blob).to(f"cuda:{device}")


def segment_ids_to_row_splits(coo: torch.Tensor, batch: int) -> torch.Tensor:
    """Device COO (int64 [nnz, stride]) -> int32 row_splits[batch+1]."""
    assert coo.dtype == torch.int64 and coo.is_cuda and coo.dim() == 2
    rs = torch.empty(batch + 1, dtype=torch.int32, device=coo.device)
    s = torch.cuda.current_stream(coo.device)
    L.check(L.lib.recom_segment_ids_to_row_splits(coo.data_ptr() if coo.numel() else None,
                                                  coo.shape[0], coo.shape[1], batch,
                                                  rs.data_ptr(), s.cuda_stream))
    return rs


def fingerprint64(bytes_dev: torch.Tensor, offsets_dev: torch.Tensor) -> torch.Tensor:
    n = offsets_dev.numel() - 1
    out = torch.empty(n, dtype=torch.int64, device=offsets_dev.device)
    s = torch.cuda.current_stream(offsets_dev.device)
    L.check(L.lib.recom_fingerprint64(bytes_dev.data_ptr(), offsets_dev.data_ptr(), n,
                                      out.data_ptr(), s.cuda_stream))
    return out


def string_to_hash_bucket_fast(bytes_dev: torch.Tensor, offsets_dev: torch.Tensor,
                               num_buckets: int) -> torch.Tensor:
    n = offsets_dev.numel() - 1
    out = torch.empty(n, dtype=torch.int64, device=offsets_dev.device)
    s = torch.cuda.current_stream(offsets_dev.device)
    L.check(L.lib.recom_string_to_hash_bucket_fast(bytes_dev.data_ptr(), offsets_dev.data_ptr(), n,
                                                   num_buckets, out.data_ptr(), s.cuda_stream))
    return out


def dense_relu(a: torch.Tensor, w_nk: torch.Tensor, bias: Optional[torch.Tensor] = None,
               relu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = relu(a @ w_nk.T + bias) on the tcgen05 tensor cores (recom_dense_relu,
    csrc/dense_layer.cu). a: bf16 [M, K] (row stride a.stride(0): the concat tensor of a bf16 plan),
    w_nk: bf16 [N, K] contiguous (the layer's [K, N] kernel transposed once), bias: fp32 [N] or None.
    -> fp32 [M, N]."""
    assert a.is_cuda and a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1
    assert w_nk.is_cuda and w_nk.dtype == torch.bfloat16 and w_nk.is_contiguous() and w_nk.shape[1] == a.shape[1]
    m, k = a.shape
    n = w_nk.shape[0]
    if bias is not None:
        assert bias.is_cuda and bias.dtype == torch.float32 and bias.numel() == n and bias.is_contiguous()
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (m, n)
    s = torch.cuda.current_stream(a.device)
    L.check(L.lib.recom_dense_relu(a.data_ptr(), m, k, a.stride(0), w_nk.data_ptr(), n,
                                   bias.data_ptr() if bias is not None else None, 1 if relu else 0,
                                   out.data_ptr(), s.cuda_stream))
    return out
