#!/usr/bin/env python
"""Summarise a per-warp timer dump (make prof build; RECOM_B200_PROF_DUMP).

Rows = warps (CTA-major); columns: cycles in [claims + run entry, FAST phase 1, FAST phase 2,
wait for the transform tasks, gather tiles, transform tasks], then the warp's start / end in
globaltimer ns."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.float64)
a = a[a[:, 6] > 0]
CYC = 1.0 / 1.965  # ns per SM cycle at 1965 MHz
t0 = a[:, 6].min()
dur = (a[:, 7] - a[:, 6]) / 1e3
print(f"warps {a.shape[0]}  kernel span {(a[:,7].max()-t0)/1e3:.1f} us  start spread {(a[:,6].max()-t0)/1e3:.1f} us")
names = ["claim/enter", "fast_p1", "fast_p2", "xform_wait", "gather", "xform"]
dense = a[:, 1] > 0
sparse = a[:, 1] == 0
for nm, m in (("fast", dense), ("queue", sparse)):
    if not m.any():
        continue
    d = dur[m]
    end = (a[m, 7] - t0) / 1e3
    print(f"{nm}: warps {m.sum()}  busy mean {d.mean():.1f} p50 {np.median(d):.1f} p90 {np.percentile(d,90):.1f} max {d.max():.1f} us; "
          f"end mean {end.mean():.1f} p90 {np.percentile(end,90):.1f} max {end.max():.1f} us")
    for i, n in enumerate(names):
        c = a[m, i] * CYC / 1e3
        if c.sum() > 0:
            print(f"   {n:<9} mean {c.mean():6.2f} p50 {np.median(c):6.2f} p90 {np.percentile(c,90):6.2f} max {c.max():6.2f} us")
# the slowest dense warps
idx = np.argsort(-(a[:, 7] - t0))[:8]
for i in idx:
    print(f"  late warp row {i} (cta {i//16} warp {i%16}): end {(a[i,7]-t0)/1e3:.1f} us, start {(a[i,6]-t0)/1e3:.1f} phases " + " ".join(f"{a[i,j]*CYC/1e3:.1f}" for j in range(6)))
