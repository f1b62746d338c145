"""Multi-GPU partitioning of the fused embedding-column stage.

The reference has no multi-GPU path (one process, one GPU, CPU spill for tables above
256 MiB: cuda_emitter.cc:1080-1094). Columns are independent units, so the path shards
with exactly one exchange step at the end:

  Mode A  column-group sharding (`partition_columns`): every rank owns a CONTIGUOUS run of
          the concat row (so rank order == concat order), holds only those tables, computes
          out_g[B, W_g]; ranks exchange the slices. Used when the tables do not fit one GPU
          or there are too many columns.
  Mode B  table row-sharding (`RowShard`): rows r of a table live on rank r % G; every rank
          reduces its local rows into partial sums [B, D] (+ counts), combined with
          reduce-scatter / all-reduce; mean / sqrtn scaling is applied after the reduction.
  Mode C  replicas: everything fits (models E, F): batch split, no collective.

This module is host logic plus torch.distributed plumbing (one process per GPU). The
exchange of Mode A is an NCCL all-to-all of row blocks (every rank keeps its B/G rows of the
concat tensor: what a data-parallel consumer needs, and the exchanged bytes per rank shrink with
G) or an all-gather of the padded slices (every rank gets all B rows), each followed by one
concatenation copy; the kernel-side variant (stores straight into peer concat tensors over
NVLink) is described in DESIGN.md.
"""
from __future__ import annotations

import os

import copy
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .schema import Column, Feature, Model


# ---------------------------------------------------------------------------------------
# Mode A: contiguous column groups
# ---------------------------------------------------------------------------------------
@dataclass
class ColumnShard:
    rank: int
    first_col: int          # [first_col, last_col) of the global model, in concat order
    last_col: int
    out_begin: int          # element span of this shard in the global concat row
    out_end: int
    model: Model            # local model: local table indices, local slots, offsets from 0
    table_ids: List[int]    # global table index of every local table

    @property
    def width(self) -> int:
        return self.out_end - self.out_begin


def column_cost(c: Column, es: int, ids_per_row: float = 5.5) -> float:
    """Relative cost of a column per sample: bytes written + bytes gathered (multi-hot rows
    gather ~ids_per_row table rows; latency-bound HBM gathers weigh more)."""
    out_bytes = c.dim * es
    if c.lookup == L.LOOKUP_DENSE:
        return out_bytes + (2.0 * out_bytes if c.transform in (L.XFORM_HASH_STR, L.XFORM_HASH_INT) else 0.0)
    big = c.table_rows * c.dim * es > (32 << 20)
    return out_bytes + ids_per_row * out_bytes * (4.0 if big else 2.0)


def partition_columns(model: Model, world: int, max_table_bytes: Optional[int] = None) -> List[ColumnShard]:
    """Split the concat row into `world` contiguous runs of columns with balanced cost,
    optionally bounded by table bytes per rank. Deterministic: every rank computes the same
    partition from the same model."""
    assert world >= 1
    if world > len(model.columns):
        raise ValueError(f"{len(model.columns)} columns cannot be split over {world} ranks "
                         "(every rank needs at least one column)")
    es = 2 if model.out_dtype == L.BF16 else 4
    order = sorted(range(len(model.columns)), key=lambda i: model.columns[i].out_offset)
    assert order == list(range(len(model.columns))), "model columns must be in concat order"
    costs = np.array([column_cost(c, es) for c in model.columns])
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    cuts = [0]
    for g in range(1, world):
        target = cum[-1] * g / world
        k = int(np.searchsorted(cum, target, side="left"))
        # at least one column for this rank and for each of the ranks still to come
        k = min(max(k, cuts[-1] + 1), len(model.columns) - (world - g))
        cuts.append(k)
    cuts.append(len(model.columns))
    shards = []
    for g in range(world):
        lo, hi = cuts[g], cuts[g + 1]
        cols = [copy.copy(c) for c in model.columns[lo:hi]]
        table_ids: List[int] = []
        remap: Dict[int, int] = {}
        for c in cols:
            if c.table not in remap:
                remap[c.table] = len(table_ids)
                table_ids.append(c.table)
            c.table = remap[c.table]
        local = Model(cols, [model.table_shapes[t] for t in table_ids], out_dtype=model.out_dtype,
                      name=f"{model.name}.shard{g}of{world}")
        if cols:
            local.finalize()
        out_begin = model.columns[lo].out_offset if lo < hi else (shards[-1].out_end if shards else 0)
        out_end = out_begin + sum(c.dim for c in cols)
        sh = ColumnShard(g, lo, hi, out_begin, out_end, local, table_ids)
        if max_table_bytes is not None:
            assert local.table_bytes() <= max_table_bytes, (
                f"shard {g}: {local.table_bytes()} table bytes exceed {max_table_bytes}")
        shards.append(sh)
    return shards


def shard_features(shard: ColumnShard, features: Dict[str, Feature]) -> Dict[str, Feature]:
    """Inputs are replicated (ids are small next to tables): a rank keeps what its columns read."""
    keys = {c.key for c in shard.model.columns}
    return {k: v for k, v in features.items() if k in keys}


def assemble_host(shards: Sequence[ColumnShard], outs: Sequence[np.ndarray]) -> np.ndarray:
    """Host "fake all-gather": concatenate shard outputs in rank (== concat) order."""
    return np.concatenate([o[:, :s.width] for s, o in zip(shards, outs)], axis=1)


def all_gather_columns(local_out, shards: Sequence[ColumnShard], group=None):
    """torch.distributed exchange of Mode A: all-gather of the (padded) slices, then one
    concatenation. local_out: [B, W_rank] tensor. Works with nccl (GPU) and gloo (CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    wmax = max(s.width for s in shards)
    B = local_out.shape[0]
    if local_out.shape[1] == wmax:
        padded = local_out.contiguous()
    else:
        padded = torch.zeros((B, wmax), dtype=local_out.dtype, device=local_out.device)
        padded[:, :local_out.shape[1]] = local_out
    flat = torch.empty((world * B, wmax), dtype=local_out.dtype, device=local_out.device)
    dist.all_gather_into_tensor(flat, padded, group=group)
    gathered = flat.view(world, B, wmax)
    return torch.cat([gathered[g, :, :shards[g].width] for g in range(world)], dim=1)


def rows_per_rank(batch: int, world: int) -> int:
    """Row ownership of every Mode-A exchange (all-to-all and the fused peer stores alike):
    blocks of ceil(batch / world) rows, the last ranks may own fewer (or none)."""
    return (batch + world - 1) // world


def batch_rows(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the batch that rank `rank` owns after the exchange."""
    rpr = rows_per_rank(batch, world)
    return min(batch, rank * rpr), min(batch, (rank + 1) * rpr)


def all_to_all_columns(local_out, shards: Sequence[ColumnShard], group=None):
    """Exchange of Mode A when the consumer is data parallel (the dense layers of the model
    split the batch over the same ranks): every rank ends with ITS rows of the full concat
    tensor, [B_rank, sum D], instead of all B rows. One all-to-all of row blocks of the local
    slice - a rank receives (B / G) * sum D elements instead of B * sum D, so the exchanged
    bytes per rank shrink with the number of ranks - then one concatenation of the received
    [B_rank, W_g] blocks. Works with nccl (GPU) and gloo (CPU tests)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    B, w_me = local_out.shape
    assert w_me == shards[rank].width
    lo, hi = batch_rows(B, world, rank)
    rows_me = hi - lo
    send = local_out.contiguous().reshape(-1)
    in_splits = [(batch_rows(B, world, d)[1] - batch_rows(B, world, d)[0]) * w_me for d in range(world)]
    out_splits = [rows_me * shards[g].width for g in range(world)]
    recv = torch.empty(sum(out_splits), dtype=local_out.dtype, device=local_out.device)
    dist.all_to_all_single(recv, send, out_splits, in_splits, group=group)
    blocks, o = [], 0
    for g in range(world):
        blocks.append(recv[o:o + out_splits[g]].view(rows_me, shards[g].width))
        o += out_splits[g]
    return torch.cat(blocks, dim=1)


class PeerOutputs:
    """The concat tensors of a fused Mode-A exchange: every rank owns [rows_per_rank, sum D]
    (its rows of the batch) in a cudaMalloc'ed buffer that all ranks of the node map through
    CUDA IPC, so that the kernel of every rank stores its column slice of every row straight
    into the row's owner over NVLink (recom_plan_run_sharded). One process per GPU."""

    def __init__(self, shards: Sequence[ColumnShard], batch: int, out_width: int, elem_bytes: int,
                 group=None):
        import ctypes as C

        import torch.distributed as dist

        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows_per_rank = rows_per_rank(batch, self.world)
        if self.rows_per_rank < 32:
            raise ValueError(f"fused exchange needs >= 32 rows per rank (batch {batch} over "
                             f"{self.world} ranks gives {self.rows_per_rank}): use all_to_all_columns")
        if len(shards) != self.world or any(s.last_col <= s.first_col for s in shards):
            raise ValueError("every rank needs a non-empty column shard")
        self.out_width, self.elem_bytes = out_width, elem_bytes
        self.nbytes = self.rows_per_rank * out_width * elem_bytes
        # the completion flags (recom_b200.h, RECOM_SHARD_FLAG_WORDS) follow the tensor
        self.flags_off = (self.nbytes + 255) // 256 * 256
        local = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        L.check(L.lib.recom_ipc_alloc(self.flags_off + 4 * L.SHARD_FLAG_WORDS, C.byref(local), handle))
        self.local_ptr = local.value
        self.epoch = 0
        self._flags_view().zero_()
        import torch
        torch.cuda.synchronize()
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self._opened = []
        self.ptrs = np.zeros(self.world, dtype=np.uint64)
        for g in range(self.world):
            if g == self.rank:
                self.ptrs[g] = self.local_ptr
                continue
            buf = (C.c_uint8 * 64).from_buffer_copy(handles[g])
            p = C.c_void_p()
            L.check(L.lib.recom_ipc_open(buf, C.byref(p)))
            self._opened.append(p.value)
            self.ptrs[g] = p.value
        sh = shards[self.rank]
        self.flag_ptrs = self.ptrs + np.uint64(self.flags_off)
        self.shard_out = L.ShardOut(world=self.world, col_offset=sh.out_begin, global_width=out_width,
                                    rank=self.rank, rows_per_rank=self.rows_per_rank,
                                    peer_out=self.ptrs.ctypes.data,
                                    peer_flags=self.flag_ptrs.ctypes.data, epoch=0, flags=0)
        dist.barrier(group=group)  # every rank's flags are zeroed before anyone signals

    def _flags_view(self):
        import torch

        class _Arr:
            pass

        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (L.SHARD_FLAG_WORDS,), "typestr": "<i4",
                                      "data": (self.local_ptr + self.flags_off, False), "version": 2}
        return torch.as_tensor(a, device=f"cuda:{torch.cuda.current_device()}")

    def next_step(self, wait_in_kernel: bool = False) -> "L.ShardOut":
        """The argument of this step's recom_plan_run_sharded: the epoch moves on by one.
        wait_in_kernel: the launch itself ends only when all ranks' slices of this rank's rows
        have landed (RECOM_SHARD_WAIT_IN_KERNEL), no wait() needed."""
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF
        self.shard_out.epoch = self.epoch
        self.shard_out.flags = L.SHARD_WAIT_IN_KERNEL if wait_in_kernel else 0
        return self.shard_out

    def wait(self, stream=None, timeout_ms: int = 0) -> None:
        """Order `stream` after the kernels of this step on all ranks (recom_shard_wait)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream()
        L.check(L.lib.recom_shard_wait(self.local_ptr + self.flags_off, self.world, self.epoch,
                                       timeout_ms, s.cuda_stream))

    def timed_out(self) -> bool:
        return bool(self._flags_view()[L.SHARD_FLAG_WORDS - 1].item())

    def local_tensor(self, dtype):
        """This rank's [rows_per_rank, sum D] tensor as a torch view of the IPC buffer."""
        import torch

        class _Arr:
            pass

        a = _Arr()
        typestr = {torch.float32: "<f4", torch.bfloat16: "<u2"}[dtype]
        a.__cuda_array_interface__ = {"shape": (self.rows_per_rank, self.out_width), "typestr": typestr,
                                      "data": (self.local_ptr, False), "version": 2}
        t = torch.as_tensor(a, device=f"cuda:{torch.cuda.current_device()}")
        return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t

    def close(self):
        import torch
        torch.cuda.synchronize()
        for p in self._opened:
            L.lib.recom_ipc_close(p)
        self._opened = []
        if self.local_ptr:
            L.lib.recom_ipc_free(self.local_ptr)
            self.local_ptr = None


# ---------------------------------------------------------------------------------------
# Mode B: table row-sharding
# ---------------------------------------------------------------------------------------
@dataclass
class RowShard:
    """rows r of every table live on rank r % world, at local row r // world."""

    rank: int
    world: int

    def local_rows(self, table_rows: int) -> int:
        return (table_rows - self.rank + self.world - 1) // self.world

    def local_table(self, table: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(table[self.rank::self.world])

    def local_model(self, model: Model) -> Model:
        """Same columns, partial SUMs over local rows: combiners become SUM (the scaling is
        applied after the cross-rank reduction), ids arrive pre-transformed."""
        cols = []
        for c in model.columns:
            c2 = copy.copy(c)
            c2.lookup = L.LOOKUP_SEGMENT
            c2.combiner = L.COMBINE_SUM
            c2.transform, c2.value_dtype, c2.boundaries, c2.num_buckets = L.XFORM_NONE, L.I64, None, 0
            c2.seg_format = L.SEG_ROW_SPLITS_I32
            c2.table_rows = max(self.local_rows(c.table_rows), 1)
            c2.input_key = f"{c.name}@rowshard"
            cols.append(c2)
        shapes = [(max(self.local_rows(r), 1), d) for r, d in model.table_shapes]
        return Model(cols, shapes, out_dtype=model.out_dtype, name=f"{model.name}.rows{self.rank}").finalize()

    def local_feature(self, ids: np.ndarray, row_of: np.ndarray, batch: int) -> Tuple[Feature, np.ndarray]:
        """Global row ids (already hashed / bucketized) + their sample -> the local ids this
        rank owns, as row_splits; also returns the per-sample TOTAL id count (for mean/sqrtn)."""
        mine = (ids % self.world) == self.rank
        lids = (ids[mine] // self.world).astype(np.int64)
        lrow = row_of[mine]
        counts = np.bincount(lrow, minlength=batch).astype(np.int64)
        rs = np.zeros(batch + 1, np.int32)
        rs[1:] = np.cumsum(counts)
        total = np.bincount(row_of, minlength=batch).astype(np.int64)
        return Feature(lids, row_splits=rs), total


class RowShardedStage:
    """Mode B on the device: this rank's share of a model whose TABLE ROWS are sharded over
    `world` ranks (row r on rank r % world, local row r // world). Inputs are replicated; every
    rank runs the whole fused stage over the rows it owns (the kernel skips ids of other ranks'
    rows after the hash / bucketize, plan option row_shard), the [batch, out_width] partial sums
    are reduce-scattered over the ranks (NCCL), and every rank scales the batch/world rows it keeps
    (recom_plan_scale_rows). The nearest logic in the reference is its CPU spill of tables above
    256 MiB (cuda_emitter.cc:1080-1094)."""

    def __init__(self, model: Model, world: int, rank: int, device: int = 0):
        from . import runtime as R

        self.model, self.world, self.rank, self.device = model, world, rank, device
        self.plan = R.Plan(model, device=device, row_shard=(world, rank))

    def local_tables(self, tables_np):
        """Host tables -> this rank's row shards on the device."""
        from . import runtime as R

        return R.tables_from_numpy(self.model, [np.ascontiguousarray(t[self.rank::self.world]) if t.shape[0] > self.rank
                                                else np.zeros((1, t.shape[1]), t.dtype) for t in tables_np],
                                   self.device)

    def partial(self, req, blob_dev, ptrs, out, ws, stream=None):
        """-> out: [batch, out_width] partial sums over this rank's rows."""
        self.plan.run(req, blob_dev, ptrs, out, ws, stream=stream)
        return out

    def finish(self, req, blob_dev, ptrs, ws, partial, process_group=None):
        """reduce-scatter of the partial sums over the ranks, then the scaling of the rows this
        rank keeps. -> ([rows_per_rank, out_width] tensor, first row). batch % world must be 0."""
        import torch
        import torch.distributed as dist

        B = partial.shape[0]
        assert B % self.world == 0, "Mode B keeps batch / world rows per rank"
        rpr = B // self.world
        mine = torch.empty((rpr, partial.shape[1]), dtype=partial.dtype, device=partial.device)
        if self.world > 1:
            dist.reduce_scatter_tensor(mine, partial, op=dist.ReduceOp.SUM, group=process_group)
        else:
            mine.copy_(partial)
        self.plan.scale_rows(req, blob_dev, ptrs, ws, mine, self.rank * rpr)
        return mine, self.rank * rpr


def combine_scale(combiner: int, counts: np.ndarray) -> np.ndarray:
    """Per-sample factor applied after the partial sums were reduced (TF semantics:
    mean divides by n, sqrtn by sqrt(n); empty rows stay 0)."""
    n = np.maximum(counts.astype(np.float64), 1.0)
    if combiner == L.COMBINE_MEAN:
        return (1.0 / n).astype(np.float32)
    if combiner == L.COMBINE_SQRTN:
        return (1.0 / np.sqrt(n)).astype(np.float32)
    return np.ones_like(n, dtype=np.float32)


# ---------------------------------------------------------------------------------------
# bench leg: the >180 GB large-table configuration (BASELINE.json config 4)
# ---------------------------------------------------------------------------------------
def bench_large_table(args, world: int, rank: int, local: int):
    """`bench.py --workload large_table`: prints one JSON line on rank 0."""
    import json

    import torch.distributed as dist

    line = measure_large_table(args, world, rank, local)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_large_table(args, world: int, rank: int, local: int, rows: int = 0, steps: int = 0):
    """BASELINE config 4: 1000 multi-hot mean columns over [rows, 32] fp32 tables (192 GB),
    column-group sharded over `world` GPUs, the exchange fused into the kernel (or an NCCL
    exchange). Collective over the `world` ranks of the default process group (world == 1: no
    collective is called). -> the JSON line (dict) on every rank."""
    import ctypes as C
    import json
    import time

    import torch
    import torch.distributed as dist

    from . import runtime as R
    from . import synth
    from .schema import algorithmic_bytes, pack_request

    n_cols = int(getattr(args, "lt_columns", 0) or 1000)
    rows = int(rows or getattr(args, "lt_rows", 0) or 1_500_000)
    batch = args.batch or 2048
    model = synth.large_table_model(n_cols, rows, 32)
    per_gpu_budget = 170 << 30
    if model.table_bytes() / world > per_gpu_budget:
        return {"metric": "embedding_column_stage_samples_per_sec", "unavailable":
                f"{model.table_bytes() / 1e9:.0f} GB of tables need more than {world} GPU(s)"}
    shards = partition_columns(model, world)
    sh = shards[rank]
    plan = R.Plan(sh.model, device=local)
    tables = R.make_tables(sh.model, device=local, seed=rank)
    ptrs = R.table_pointers(tables)
    K, W, S = int(steps or args.steps), max(args.warmup, 3), max(args.sets, 2)
    reqs = []
    for s in range(S):
        feats = synth.make_features(model, batch, seed=100 + s)  # replicated inputs, same on all ranks
        reqs.append(pack_request(sh.model, batch, shard_features(sh, feats)))
    blobs = [R.upload_blob(r, local) for r in reqs]
    outs = [plan.alloc_out(batch) for _ in range(S)]
    wss = [plan.alloc_workspace(batch, r.sparse_ids) for r in reqs]
    stream = torch.cuda.current_stream(local)
    alg = float(np.mean([algorithmic_bytes(sh.model, r) for r in reqs]))

    exchange = getattr(args, "lt_exchange", None) or "fused"
    if world == 1:
        exchange = "none"
    peers, flag = None, None
    if exchange == "fused":
        # rows go straight into their owners' concat tensors from the kernel; the last warp of
        # every rank's kernel signals all peers, and a one-warp wait orders the consumers
        assert S >= 2, "the fused exchange needs two rotating output sets"
        peers = [PeerOutputs(shards, batch, model.out_width, 4) for _ in range(S)]
    # The consumers' stream: the wait for step i (and whatever reads the concat rows) runs here
    # while the launch stream already computes step i + 1; a launch only waits for the
    # consumers of `lag` steps back, so the ranks are not forced into lock step by every wait
    # (a rank that has passed the wait of step j - lag may overwrite the set of step j - 2 lag
    # or older: S >= 2 lag). lag 0: wait on the launch stream itself.
    lag = min(int(getattr(args, "lt_lag", 0)), S // 2) if exchange == "fused" else 0
    wait_mode = getattr(args, "lt_wait", None) or "kernel"  # barrier inside the kernel, or a wait kernel
    if wait_mode == "kernel":
        lag = 0
    cstream = torch.cuda.Stream(local) if lag > 0 else None
    ev_k = [torch.cuda.Event() for _ in range(S)]
    ev_c = [torch.cuda.Event() for _ in range(S)]
    n_steps = [0]

    def step(i):
        s = i % S
        if exchange == "fused":
            j = n_steps[0]
            n_steps[0] += 1
            s = j % S  # the sets rotate in launch order
            if lag > 0 and j >= lag:
                stream.wait_event(ev_c[(j - lag) % S])
            plan.run_sharded(reqs[s], blobs[s], ptrs, wss[s], peers[s].next_step(wait_mode == "kernel"),
                             stream=stream)
            if wait_mode == "kernel":
                pass
            elif lag == 0:
                peers[s].wait(stream)
            else:
                ev_k[j % S].record(stream)
                cstream.wait_event(ev_k[j % S])
                peers[s].wait(cstream)
                ev_c[j % S].record(cstream)  # (a real pipeline records it after its readers)
            return peers[s].local_tensor(torch.float32)
        plan.run(reqs[s], blobs[s], ptrs, outs[s], wss[s], stream=stream)
        if world == 1:
            return outs[s]
        if exchange == "all_gather":
            return all_gather_columns(outs[s], shards)
        return all_to_all_columns(outs[s], shards)

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    for i in range(W):
        full = step(i)
    if exchange == "fused":
        want_rows = peers[0].rows_per_rank
    else:
        want_rows = batch if (world == 1 or exchange == "all_gather") else (
            batch_rows(batch, world, rank)[1] - batch_rows(batch, world, rank)[0])
    assert full.shape == (want_rows, model.out_width)
    barrier()
    if exchange == "fused":
        # the fused exchange leaves exactly what the NCCL all-to-all of the local slices leaves
        s0 = n_steps[0] % S
        plan.run(reqs[s0], blobs[s0], ptrs, outs[s0], wss[s0], stream=stream)
        ref = all_to_all_columns(outs[s0], shards)
        got = step(0)
        barrier()
        assert torch.equal(got[: ref.shape[0]], ref), "fused exchange differs from the NCCL all-to-all"
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        step(i)
    if cstream is not None:
        stream.wait_event(ev_c[(n_steps[0] - 1) % S])  # the last step's wait is inside the timed region
    e1.record(stream)
    barrier()
    if exchange == "fused":
        assert not any(p.timed_out() for p in peers), "a completion signal never arrived"
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    alg_t = torch.tensor([alg], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(alg_t, op=dist.ReduceOp.SUM)
    ms = float(t[0]) / K
    line = {
        "metric": "embedding_column_stage_samples_per_sec", "value": batch / (ms * 1e-3),
        "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"large_table_{n_cols}x{rows}x32_b{batch}", "columns": n_cols,
                   "batch": batch, "table_bytes": model.table_bytes(),
                   "parallelism": (f"column-group x{world}, exchange fused into the kernel: rows stored straight into the "
                                   f"owners' [B/{world}, sum D] tensors over NVLink (IPC peer memory); completion "
                                   f"flags stored to every peer by the kernel's last warp, " +
                                   ("which also waits for all ranks' flags: compute, exchange and barrier in one kernel"
                                    if wait_mode == "kernel" else
                                    f"one-warp wait kernel {'on the launch stream' if lag == 0 else f'on a second stream, launches run {lag} step(s) ahead'}") +
                                   " (no NCCL call in the step)" if exchange == "fused" else
                                   f"column-group x{world}, NCCL all-gather of [B, W_g] slices" if exchange == "all_gather"
                                   else f"column-group x{world}, NCCL all-to-all of [B/{world}, W_g] row blocks "
                                        f"(every rank keeps its B/{world} rows of the concat tensor)"),
                   "l2": f"rotating {S} request/output sets",
                   "library": os.path.basename(L.LIB_PATH)},
        "roofline": {"bound": "hbm", "achieved": float(alg_t[0]) / (ms * 1e-3) / 1e9,
                     "unit": "GB/s", "note": "aggregate algorithmic bytes over all ranks / step time"},
        "gpu_launches": (plan.last_launch_count() + (1 if exchange == "fused" and wait_mode != "kernel" else 0)) * K,
        "exchange": exchange,
        # the fused exchange was compared (torch.equal) with the NCCL all-to-all of the same slices
        "parity_vs_all_to_all": exchange == "fused",
    }
    if peers:
        barrier()
        for p in peers:
            p.close()
    del tables, blobs, outs, wss, plan
    torch.cuda.empty_cache()
    return line
