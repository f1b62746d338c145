#!/usr/bin/env python
"""Executed code footprint of a kernel from an .ncu-rep: SASS instructions with a non-zero execution
count, instruction-cache related metrics, stall breakdown."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:] if len(r) >= len(h)]
ie = h.index('Instructions Executed'); si = h.index('# Samples')
ex = [int(r[ie] or 0) for r in body]
n_exec = sum(1 for e in ex if e > 0)
print(f"SASS instructions {len(body)} ({len(body)*16/1024:.0f} KB), executed at least once {n_exec} ({n_exec*16/1024:.0f} KB)")
tot = sum(ex)
# footprint that covers 90 / 99 % of executed instructions
srt = sorted(ex, reverse=True); acc = 0
for frac in (0.5, 0.9, 0.99):
    a = 0; n = 0
    for e in srt:
        a += e; n += 1
        if a >= frac * tot: break
    print(f"  {int(frac*100)}% of the {tot} executed warp instructions come from {n} SASS instructions ({n*16/1024:.1f} KB)")
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
t = collections.Counter()
for r in body:
    for s in stalls: t[s] += int(r[h.index(s)] or 0)
total = sum(int(r[si] or 0) for r in body)
print("samples", total, "stalls:", ", ".join(f"{s[6:]}={v} ({100*v/max(total,1):.0f}%)" for s, v in t.most_common(9)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
for k, u, v in zip(rr[0], rr[1], rr[2]):
    if any(x in k for x in ("gcc__", "icc", "inst_executed.sum", "issue_active.avg.pct", "gpu__time_duration.sum", "warps_active.avg.pct", "registers_per_thread")) and "pct_of_peak_sustained_elapsed" in k or k in ("smsp__inst_executed.sum", "gpu__time_duration.sum", "launch__registers_per_thread"):
        print(f"  {k} [{u}] = {v}")
