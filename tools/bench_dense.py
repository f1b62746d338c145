#!/usr/bin/env python
"""Time recom_dense_relu on the first top-MLP layer of model E: A = the bf16
concat tensor [2048, 8120], W [1024, 8120]; CUDA events, L2 flushed between iterations by rotating
operand sets larger than L2; cuBLAS (torch.matmul, bf16) beside it. One JSON line."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from recom_b200 import runtime as R  # noqa: E402


def main():
    m, k, n = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (2048, 8120, 1024)))
    sets, steps = 6, 100
    g = torch.Generator(device="cuda").manual_seed(0)
    A = [torch.randn((m, k), device="cuda", generator=g).to(torch.bfloat16) for _ in range(sets)]
    W = [(torch.randn((n, k), device="cuda", generator=g) / np.sqrt(k)).to(torch.bfloat16) for _ in range(sets)]
    bias = torch.randn((n,), device="cuda", generator=g)
    out = torch.empty((m, n), device="cuda")

    def timed(fn):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i in range(steps):
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        torch.cuda.synchronize()
        t = np.array([a.elapsed_time(b) for a, b in ev])
        return float(t.mean()), float(np.percentile(t, 50))

    ours, ours50 = timed(lambda i: R.dense_relu(A[i % sets], W[i % sets], bias, True, out))
    Wt = [w.t().contiguous() for w in W]
    cub, cub50 = timed(lambda i: torch.relu(torch.addmm(bias.to(torch.bfloat16), A[i % sets], Wt[i % sets])))
    flops = 2.0 * m * n * k
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    print(json.dumps({"op": "recom_dense_relu", "m": m, "k": k, "n": n, "ms": ours, "p50_ms": ours50,
                      "tflops": flops / ours / 1e9, "cublas_bf16_ms": cub, "cublas_bf16_tflops": flops / cub / 1e9,
                      "operand_sets_mb": sets * (m * k + n * k) * 2 / 1e6, "peaks": peaks}))


if __name__ == "__main__":
    main()
