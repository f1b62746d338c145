#!/bin/bash
# batch sweep (BASELINE.json config 3) and the other configurations: one line each
line() { python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('%-28s us/step %8.2f  p50 %8.2f  p99 %8.2f  Msamples/s %8.2f  frac %.3f  e2e_us %8.1f' % (d['config']['workload'], d['ms_per_step']*1e3, d['p50_ms']*1e3, d['p99_ms']*1e3, d['value']/1e6, d['roofline']['frac'], d['e2e']['ms_per_step']*1e3))
"; }
for b in 1 16 128 256 512 1024 2048 4096 8192; do
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload model_E --batch $b 2>/dev/null | line
done
for w in model_F model_E_bf16 mixed1000 mixed1000_bf16 production2000 micro100; do
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $w 2>/dev/null | line
done
