#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []
cur_file = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        stalls = [(i, c[6:]) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
        ie = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "" or r[0] == "Function Name":
        continue
    try:
        n = int(r[si])
    except ValueError:
        continue
    if n == 0:
        continue
    st = sorted(((int(r[i] or 0), name) for i, name in stalls), reverse=True)[:3]
    out.append((n, cur_file, r[0], r[1].strip()[:90], ", ".join(f"{name}={v}" for v, name in st if v), r[ie]))
tot = sum(o[0] for o in out)
print("total samples", tot)
for n, f, ln, src, st, ie in sorted(out, reverse=True)[:topn]:
    print(f"{n:5d} {100*n/tot:5.1f}%  {f}:{ln:<5} [{st}] inst={ie}\n        {src}")
