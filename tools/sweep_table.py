#!/usr/bin/env python
"""bench.py JSON lines (tools/sweep_batches.sh) -> the markdown table of BASELINE.md section 5."""
import json, sys
print("| Config | Batch | GPUs | M samples/s | µs/step | p50 µs | p99 µs | achieved GB/s | % of measured HBM | e2e µs/step | CPU baseline samples/s (cores) | parity |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for l in open(sys.argv[1]):
    try: d = json.loads(l)
    except Exception: continue
    if "config" not in d: continue
    c, r = d["config"], d["roofline"]
    cpu = d.get("cpu_baseline")
    par = d.get("parity") or {}
    print(f"| {c['workload'].rsplit('_b', 1)[0]} | {c['batch']} | {d['n_gpus']} | {d['value']/1e6:.2f} | {d['ms_per_step']*1e3:.1f} | "
          f"{d['p50_ms']*1e3:.1f} | {d['p99_ms']*1e3:.1f} | {r['achieved']:.0f} | {100*r['frac']:.1f} | {d['e2e']['ms_per_step']*1e3:.0f} | "
          f"{(str(round(cpu['value'])) + ' (' + str(cpu['cores']) + ')') if cpu else '-'} | "
          f"{'oracle ok, max abs err %.1e' % par['max_abs_err'] if par.get('checked') else '-'} |")
