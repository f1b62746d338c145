#!/usr/bin/env python
"""bench.py — headline benchmark of the fused embedding-column stage.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--workload model_E|model_F|micro100|large_table] [--batch B]

One step = one pass of the hot path (hash/bucketize -> gather -> combine -> concat) over
one batch of synthetic features of the AE synthesized model E (1000 columns,
examples/python/dlrm.py:202 of the reference) at batch 2048, the configuration
BASELINE.json's metric is quoted on. Prints ONE JSON line (rank 0).

  value      samples/s with the request blob already resident in HBM; CUDA events on the
             launch stream around K back-to-back recom_plan_run calls, max over ranks.
  e2e        the same metric through the C-ABI call with HOST buffers: per step a pinned
             host blob -> device copy, recom_plan_run, and a device -> host read of the
             last output row, all inside the timed region.
  roofline   algorithmic bytes per launch / average step duration,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (a C restatement of the TF-CPU path the reference runs,
             OpenMP over columns) on the box's host cores, same workload, bounded sample.

N > 1: model E fits one GPU, columns and batches are independent, so ranks are replicas
(one process per GPU, own batch, no data-path collective): weak scaling. The sharded
large-table configuration is `--workload large_table` (NCCL all-gather of partial rows).

--impl reference times the reference's CPU implementation of the path. TensorFlow cannot
be installed here, so that arm is the oracle port with all host threads (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "embedding_column_stage_samples_per_sec"
UNIT = "samples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="model_E",
                    choices=["model_E", "model_F", "micro100", "large_table",
                             # diagnostic slices of model E (not headline workloads)
                             "model_E_dense", "model_E_bucketize", "model_E_sparse",
                             # other BASELINE.json configurations (informational lines)
                             "model_E_bf16", "mixed1000", "mixed1000_bf16", "production2000",
                             # config 3: variable-length multi-hot features, per-request segment counts
                             "dynamic"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--sets", type=int, default=8, help="distinct requests / rotating output buffer sets")
    ap.add_argument("--lt-columns", dest="lt_columns", type=int, default=0, help="large_table: columns")
    ap.add_argument("--lt-rows", dest="lt_rows", type=int, default=0, help="large_table: rows per table")
    ap.add_argument("--lt-exchange", dest="lt_exchange", default="fused",
                    choices=["fused", "all_to_all", "all_gather"], help="large_table: exchange of the slices")
    ap.add_argument("--lt-wait", dest="lt_wait", default="kernel", choices=["kernel", "stream"],
                    help="large_table, fused exchange: barrier inside the kernel, or a one-warp wait kernel")
    ap.add_argument("--lt-lag", dest="lt_lag", type=int, default=0,
                    help="large_table, fused exchange: steps the launch stream may run ahead of the waits")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of a timed request")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the large-table (sharded) measurement")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def build_model(workload: str, world: int = 1, rank: int = 0):
    from recom_b200 import synth

    if workload == "model_E":
        # RECOM_BENCH_LARGE_ROWS: diagnostic only (rows of the five 2^23-row tables; the line's
        # config.table_bytes shows it) - how much of a launch is the random walk over 5 GB of HBM
        if os.environ.get("RECOM_BENCH_LARGE_ROWS"):
            return synth.model_E(large_rows=int(os.environ["RECOM_BENCH_LARGE_ROWS"])), 2048
        return synth.model_E(), 2048
    if workload == "model_F":
        return synth.model_F(), 2048
    if workload == "micro100":
        return synth.microbenchmark_model(100), 256
    if workload == "model_E_dense":
        return synth.dlrm_model(880, 50, 50, 0, 0, name="E_dense"), 2048
    if workload == "model_E_bucketize":
        return synth.dlrm_model(880, 0, 0, 0, 0, name="E_bucketize"), 2048
    if workload == "model_E_sparse":
        return synth.dlrm_model(0, 0, 0, 15, 5, name="E_sparse",
                                large_rows=int(os.environ.get("RECOM_BENCH_LARGE_ROWS", synth.LARGE_ROWS))), 2048
    # informational lines (parity-tested configurations of BASELINE.json, not the headline)
    from recom_b200 import consts as L
    if workload == "model_E_bf16":
        return synth.model_E(out_dtype=L.BF16), 2048
    if workload == "mixed1000":      # config 2 extension: dims 8-64, sum / mean / sqrtn
        return synth.mixed_model(1000, seed=0), 2048
    if workload == "mixed1000_bf16":
        return synth.mixed_model(1000, seed=0, out_dtype=L.BF16), 2048
    if workload == "dynamic":  # config 3: 20 multi-hot columns (hashed strings) among 70, long tails
        return synth.dlrm_model(40, 5, 5, 15, 5, large_rows=1 << 20, name="dyn"), 2048
    if workload == "production2000":  # config 5: shared tables, duplicated (input, transform) pairs
        return synth.production_model(2000, share=4, seed=0), 2048
    raise ValueError(workload)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"  # /opt/skills/guides/B200_PROFILING.md


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_tables(model, seed=0):
    """Random-init fp32 tables in host memory (same distribution as the device tables)."""
    import torch

    g = torch.Generator()
    g.manual_seed(seed + 12345)
    out = []
    for rows, dim in model.table_shapes:
        t = torch.empty((rows, dim), dtype=torch.float32)
        t.normal_(0.0, 1.0 / float(np.sqrt(dim)), generator=g)
        out.append(t.numpy())
    return out


def time_oracle(model, batch, n_sets, seconds, warmup, steps=None):
    """Times the CPU oracle on `model` at `batch`. -> dict(samples_per_s, p50_ms, p99_ms, ...)."""
    from oracle import oracle as O
    from recom_b200 import synth
    from recom_b200.schema import pack_request

    descs, bnd = model.descs()
    tabs = host_tables(model)
    ptrs = np.array([t.ctypes.data for t in tabs], dtype=np.uint64)
    fkw = dict(dist="zipf", empty_frac=0.2) if model.name == "dyn" else {}
    reqs = [pack_request(model, batch, synth.make_features(model, batch, seed=s, **fkw))
            for s in range(n_sets)]
    out = np.zeros((batch, model.out_width), np.float32)
    # all the host cores this process may use; torchrun exports OMP_NUM_THREADS=1 to its
    # workers, which must not shrink the CPU arm (only rank 0 runs it)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or O.num_threads()
    threads = max(threads, O.num_threads())

    def one(r):
        rc = O.lib.oracle_run_model(descs.ctypes.data, len(descs), bnd.ctypes.data if bnd.size else None,
                                    model.out_width, model.out_dtype, batch, r.blob.ctypes.data,
                                    r.slot_offsets.ctypes.data, r.slot_elems.ctypes.data,
                                    ptrs.ctypes.data, out.ctypes.data, None, threads)
        assert rc == 0

    for i in range(max(warmup, 1)):
        one(reqs[i % n_sets])
    lat = []
    t0 = time.perf_counter()
    i = 0
    while True:
        t1 = time.perf_counter()
        one(reqs[i % n_sets])
        t2 = time.perf_counter()
        lat.append(t2 - t1)
        i += 1
        if steps is not None:
            if i >= steps:
                break
        elif t2 - t0 >= seconds:
            break
    total = time.perf_counter() - t0
    lat_ms = np.array(lat) * 1e3
    return {"samples_per_s": batch * len(lat) / total, "steps": len(lat), "cores": threads,
            "ms_per_step": total / len(lat) * 1e3, "p50_ms": float(np.percentile(lat_ms, 50)),
            "p99_ms": float(np.percentile(lat_ms, 99)),
            "rot_mb": (sum(r.blob.nbytes for r in reqs) + n_sets * out.nbytes) / 1e6}


def _package_without_library():
    """The reference arm needs the synthetic models and the request packer (pure numpy) but must
    not map the product's shared library: register the package without running its __init__."""
    import types

    if "recom_b200" not in sys.modules:
        pkg = types.ModuleType("recom_b200")
        pkg.__path__ = [os.path.join(ROOT, "recom_b200")]
        sys.modules["recom_b200"] = pkg


def config_of(workload, model, batch, world, rot_mb=None, sets=None):
    """The same keys on both arms (the driver compares the two `config` dicts)."""
    return {"workload": f"{workload}_b{batch}", "columns": len(model.columns), "batch": batch,
            "out_width": model.out_width, "table_bytes": model.table_bytes(),
            "parallelism": f"replicas x{world}",
            "l2": f"rotating {sets} distinct requests / output sets = {rot_mb:.0f} MB > 126 MB L2",
            "hashing": "fused (Fingerprint64 of the raw strings / integers inside the stage)",
            "segments": "int32 row_splits"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _package_without_library()
    model, batch = build_model(args.workload)
    batch = args.batch or batch
    r = time_oracle(model, batch, max(args.sets, 2), 0.0, args.warmup, steps=args.steps)
    sample = f"{r['steps']} batches of {batch} samples of {args.workload}, all {len(model.columns)} columns"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["samples_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "p50_ms": r["p50_ms"], "p99_ms": r["p99_ms"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload, model, batch, 1, r["rot_mb"], max(args.sets, 2)),
        "cpu_baseline": {"value": r["samples_per_s"], "unit": UNIT, "cores": r["cores"],
                         "kind": "port", "sample": sample},
        "e2e": {"value": r["samples_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "note": "reference arm = CPU oracle port (TensorFlow 2.6 is not installable here): "
                "un-fused per-column TF-CPU kernels restated in C, OpenMP over columns",
    }
    print(json.dumps(line), flush=True)


def check_parity(model, batch, req, tables, out, launch, stream):
    """Runs `req` on the device and through the CPU oracle (test infrastructure, the checker only)
    on the same tables; embeddings within north_star's tolerance (1e-5 relative for fp32, 1e-2 for
    bf16; against the summed magnitudes for multi-hot rows), copies bit-exact. -> summary dict."""
    import torch

    from oracle import oracle as O
    from recom_b200 import consts as K

    bf16 = model.out_dtype == K.BF16
    with torch.cuda.stream(stream):
        launch()
    stream.synchronize()
    got = out.float().cpu().numpy()
    if bf16:
        host = [t.view(torch.int16).cpu().numpy().view(np.uint16) for t in tables]
    else:
        host = [t.cpu().numpy() for t in tables]
    descs, bnd = model.descs()
    ref, _ = O.run_model(descs, bnd, model.out_width, model.out_dtype, batch, req.blob, req.slot_offsets,
                         req.slot_elems, host, req.total_ids, O.num_threads())
    rtol, atol = (1e-2, 1e-5) if bf16 else (1e-5, 1e-6)
    bad = ~np.isclose(got, ref, rtol=rtol, atol=atol)
    exact_cols = 0
    for c in model.columns:
        if c.lookup != K.LOOKUP_SEGMENT:
            sl = slice(c.out_offset, c.out_offset + c.dim)
            assert np.array_equal(got[:, sl], ref[:, sl]), f"column {c.name}: copy is not bit-exact"
            exact_cols += 1
    criterion = "numpy isclose (|err| <= atol + rtol |ref|)"
    if bad.any():
        # rows of hundreds of ids are summed in another order than the oracle's: the error of an
        # fp32 sum is bounded by the summed magnitudes, not by the (possibly cancelling) result
        absd = [np.abs(t) if not bf16 else O.f32_to_bf16_bits(np.abs((t.astype(np.uint32) << 16).view(np.float32)))
                for t in host]
        scale, _ = O.run_model(descs, bnd, model.out_width, model.out_dtype, batch, req.blob, req.slot_offsets,
                               req.slot_elems, absd, req.total_ids, O.num_threads())
        worst = float((np.abs(got - ref) / (rtol * scale + atol)).max())
        assert worst <= 1.0, f"max err / (rtol * sum|terms| + atol) = {worst:.3g}"
        criterion = f"forward bound (|err| <= atol + rtol * sum|terms|; {int(bad.sum())} elements outside plain isclose)"
    return {"checked": True, "against": "CPU oracle (oracle/recom_oracle.c)", "rtol": rtol, "atol": atol,
            "criterion": criterion,
            "max_abs_err": float(np.abs(got - ref).max()), "bit_exact_columns": exact_cols,
            "columns": len(model.columns)}


def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from recom_b200 import _lib as L
    from recom_b200 import runtime as R
    from recom_b200 import synth
    from recom_b200.schema import algorithmic_bytes, pack_request

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))

    if args.workload == "large_table":
        from recom_b200 import sharding
        return sharding.bench_large_table(args, world, rank, local)

    model, batch = build_model(args.workload)
    batch = args.batch or batch
    # S distinct requests (own ids, own row lengths, hence own slot offsets each): nothing in the
    # timed loop is served from a cache of earlier requests - the argument data of a request
    # travels inside its blob
    K, W, S = args.steps, max(args.warmup, 3), max(args.sets, 2)

    plan = R.Plan(model, device=local)
    tables = R.make_tables(model, device=local, seed=0)
    ptrs = R.table_pointers(tables)
    fkw = dict(dist="zipf", empty_frac=0.2) if args.workload == "dynamic" else {}
    reqs = [pack_request(model, batch, synth.make_features(model, batch, seed=1000 * rank + s, **fkw))
            for s in range(S)]  # (the reference arm packs the same S requests of rank 0)
    blobs = [R.upload_blob(r, local) for r in reqs]
    outs = [plan.alloc_out(batch) for _ in range(S)]
    wss = [plan.alloc_workspace(batch, r.sparse_ids) for r in reqs]
    stream = torch.cuda.Stream(device=local)
    creqs = [plan._request(reqs[s], blobs[s], ptrs, outs[s], wss[s], None, None, stream) for s in range(S)]
    alg_bytes = [plan.algorithmic_bytes(reqs[s], blobs[s], ptrs, outs[s], wss[s]) for s in range(S)]
    assert alg_bytes[0] == algorithmic_bytes(model, reqs[0])
    run = L.lib.recom_plan_run
    h = plan._h
    refs = [C.byref(c) for c in creqs]

    def barrier():
        torch.cuda.synchronize(local)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(local)

    # rotating sets: S x (blob + [B, out_width] output) exceeds the 126 MB L2
    rot_bytes = sum(b.numel() for b in blobs) + sum(o.numel() * o.element_size() for o in outs)

    with torch.cuda.stream(stream):
        for i in range(W):
            L.check(run(h, refs[i % S]))
    launches_per_step = plan.last_launch_count()
    barrier()

    # ---- parity of what is timed: one of the timed requests against the CPU oracle ----
    parity = None
    if rank == 0 and not args.no_parity:
        parity = check_parity(model, batch, reqs[0], tables, outs[0], lambda: L.check(run(h, refs[0])), stream)
        torch.cuda.synchronize(local)

    # ---- device-resident timing: K back-to-back steps -----------------------------
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # cudaProfilerStart/Stop bracket the timed region: `ncu --profile-from-start off` lists
    # exactly the launches of the K timed steps (profiles/*_launches.csv)
    torch.cuda.profiler.start()
    e0.record(stream)
    for i in range(K):
        st = run(h, refs[i % S])
        if st:
            L.check(st)
    e1.record(stream)
    barrier()
    torch.cuda.profiler.stop()
    total_ms = e0.elapsed_time(e1)

    # the same K steps with every launch fully ordered behind the previous one (no programmatic
    # dependent launch between consecutive requests): what one request costs on an idle GPU
    plan.set_request_overlap(False)
    for i in range(3):
        L.check(run(h, refs[i % S]))
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for i in range(K):
        st = run(h, refs[i % S])
        if st:
            L.check(st)
    g1.record(stream)
    barrier()
    serialized_ms = g0.elapsed_time(g1) / K
    plan.set_request_overlap(True)

    # per-step latency distribution (event pair per step)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for i in range(K):
        evs[i][0].record(stream)
        L.check(run(h, refs[i % S]))
        evs[i][1].record(stream)
    barrier()
    lat = np.array([a.elapsed_time(b) for a, b in evs])

    # ---- end to end: host blob -> H2D -> run -> D2H of the last output row ---------
    pinned = [torch.from_numpy(r.blob).pin_memory() for r in reqs]
    last_row = torch.empty(model.out_width, dtype=outs[0].dtype).pin_memory()
    h2d = int(np.mean([r.blob.nbytes for r in reqs]))
    d2h = last_row.numel() * last_row.element_size()

    # Two streams, request sets alternating between them: the H2D copy of request i+1 overlaps
    # the kernel of request i (what a serving loop does). Every step still pays its own H2D
    # copy of the blob and D2H read of the result inside the timed region.
    streams2 = [stream, torch.cuda.Stream(device=local)]
    last_rows = [last_row, torch.empty_like(last_row).pin_memory()]
    creqs2 = [plan._request(reqs[s], blobs[s], ptrs, outs[s], wss[s], None, None, streams2[s % 2])
              for s in range(S)]
    refs2 = [C.byref(c) for c in creqs2]

    def e2e_step(i):
        s = i % S
        with torch.cuda.stream(streams2[s % 2]):
            blobs[s].copy_(pinned[s], non_blocking=True)
            st = run(h, refs2[s])
            if st:
                L.check(st)
            last_rows[s % 2].copy_(outs[s][batch - 1], non_blocking=True)

    for i in range(W):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Event()
    t0 = time.perf_counter()
    f0.record(streams2[0])
    streams2[1].wait_event(f0)
    for i in range(K):
        e2e_step(i)
    join.record(streams2[1])
    streams2[0].wait_event(join)
    f1.record(streams2[0])
    streams2[0].synchronize()
    streams2[1].synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), wall_ms)
    clocks = sampler.stop()

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = world * batch * K / (total_ms * 1e-3)
    e2e_value = world * batch * K / (e2e_ms * 1e-3)

    peak, peak_src = peaks()
    mean_alg = float(np.mean(alg_bytes))
    achieved = mean_alg / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "p50_ms": float(np.percentile(lat, 50)),
        "p99_ms": float(np.percentile(lat, 99)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if outs[0].element_size() == 2 else "f32", "data": "synthetic",
        "config": config_of(args.workload, model, batch, world, rot_bytes / 1e6, S),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": mean_alg,
                     # one launch per step: the fused persistent kernel; two: the flat kernel of the
                     # column tiles + the persistent kernel of the FAST runs (capi.cu prefer_flat) -
                     # `achieved` is then the step's bytes over the step's duration (both launches)
                     "kernel": "fused_columns_kernel" if launches_per_step == 1 and not os.environ.get("RECOM_B200_FLAT")
                               else f"{launches_per_step} launches per step (column_tiles_kernel / fused_columns_kernel)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K},
        "gpu_launches": launches_per_step * K,
        # ms_per_step / value / roofline: K back-to-back requests on one stream, consecutive fused launches under
        # programmatic dependent launch (the prologue of request i+1 - records, contexts, TMA staging, blob
        # prefetch: reads only - overlaps the tail of request i; nothing of i+1 is stored before i is complete).
        # p50 / p99: one request at a time (an event pair around each launch orders it fully).
        "serialized": {"ms_per_step": serialized_ms, "frac": mean_alg / (serialized_ms * 1e-3) / 1e9 / peak,
                       "note": "same K steps with recom_plan_set_request_overlap(0): every launch starts after the previous one has ended"},
        "clocks": clocks,
        "parity_checked": bool(parity and parity["checked"]), "parity": parity,
        "requests": f"{S} distinct requests (own ids and row lengths each); per-request argument data "
                    "travels inside the blob: no argument cache, no host->device copy besides the blob",
        "e2e_note": "d2h = the last output row (the consumer of the concat tensor is the GPU MLP); "
                    "h2d = the whole request blob from pinned memory, every step",
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del tables
        torch.cuda.empty_cache()
        r = time_oracle(model, batch, 2, args.cpu_seconds, 2)
        line["cpu_baseline"] = {
            "value": r["samples_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "p50_ms": r["p50_ms"], "p99_ms": r["p99_ms"],
            "sample": f"{r['steps']} batches of {batch} samples of {args.workload} "
                      f"(all {len(model.columns)} columns), {args.cpu_seconds:.0f} s of CPU work",
        }
    if world > 1 and not args.no_sharded:
        # The configuration that needs more than one GPU (BASELINE config 4: 1000 multi-hot columns
        # over 192 GB of tables), column-group sharded over the same ranks with the exchange fused
        # into the kernel; its one-GPU equivalent (tables cut to fit one GPU) is measured first,
        # on rank 0, in this very run. Same library as the headline.
        from recom_b200 import sharding

        del tables, blobs, outs, wss, plan
        torch.cuda.empty_cache()
        lt_steps = min(K, 100)
        one = None
        if rank == 0:
            one = sharding.measure_large_table(args, 1, 0, local, rows=1_300_000, steps=lt_steps)
        barrier()
        sh = sharding.measure_large_table(args, world, rank, local, steps=lt_steps)
        if rank == 0:
            line["sharded"] = {
                "workload": sh.get("config", {}).get("workload"), "ms_per_step": sh.get("ms_per_step"),
                "value": sh.get("value"), "unit": UNIT, "n_gpus": world, "scaling": "strong",
                "exchange": sh.get("exchange"), "parity_vs_all_to_all": sh.get("parity_vs_all_to_all"),
                "one_gpu_equiv_ms_per_step": one.get("ms_per_step"),
                "one_gpu_equiv": one.get("config", {}).get("workload"),
                "speedup_vs_1gpu_equiv": (one["ms_per_step"] / sh["ms_per_step"]
                                          if one.get("ms_per_step") and sh.get("ms_per_step") else None),
                "library": sh.get("config", {}).get("library"), "unavailable": sh.get("unavailable"),
            }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
