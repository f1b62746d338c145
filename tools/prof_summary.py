#!/usr/bin/env python
"""Summarise a phase-timer dump (make prof build; RECOM_B200_PROF_DUMP)."""
import sys
import numpy as np
a = np.loadtxt(sys.argv[1], dtype=np.uint64).astype(np.float64)
names = ["looptop", "enter+sync", "resolve", "cp+issue", "-", "-", "dense_tile", "end_sync",
         "s_phase1", "s_barrier", "s_phase2", "tiles", "stagings", "prologue", "t_start", "t_end"]
CYC = 1.0 / 1.965  # ns per SM cycle at 1965 MHz
t0 = a[:, 14].min()
dur = a[:, 15] - a[:, 14]
print(f"CTAs {a.shape[0]}  kernel span {(a[:,15].max()-t0)/1e3:.1f} us  start spread {(a[:,14].max()-t0)/1e3:.1f} us  "
      f"CTA busy mean {dur.mean()/1e3:.1f} max {dur.max()/1e3:.1f} min {dur.min()/1e3:.1f} us")
print(f"tiles/CTA mean {a[:,11].mean():.2f} max {a[:,11].max():.0f}; stagings/CTA mean {a[:,12].mean():.2f}")
for i in list(range(11)) + [13]:
    c = a[:, i] * CYC
    print(f"  {names[i]:<12} mean/CTA {c.mean()/1e3:7.2f} us   per tile {c.sum()/max(a[:,11].sum(),1)/1e3:6.3f} us   max/CTA {c.max()/1e3:7.2f}")
print(f"  sum of phases mean/CTA {(a[:, list(range(11)) + [13]].sum(axis=1) * CYC).mean()/1e3:.2f} us")
