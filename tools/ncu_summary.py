#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics, stall reasons, hottest source lines."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'smsp__cycles_active.avg',
        'sm__cycles_elapsed.max', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__warps_eligible.avg.per_cycle_active']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} [{units[i]}] = {[r[i] for r in rows[2:]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
body = [r for r in rows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(rows))] if len(r) >= len(h)]
si = h.index('# Samples')
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
tot = collections.Counter()
for r in body:
    for s in stalls:
        tot[s] += int(r[h.index(s)] or 0)
total = sum(int(r[si] or 0) for r in body)
print("total samples", total, " instructions", len(body))
print("stalls:", ", ".join(f"{s[6:]}={v}" for s, v in tot.most_common(8)))
print("--- hottest SASS (samples, top stall, instr) with context")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:topn]
for i in idx:
    r = body[i]
    st = max(stalls, key=lambda s: int(r[h.index(s)] or 0))
    print(f"{r[si]:>6} {st[6:]:<12} #{i:<6} {r[1][:100]}")
