#!/bin/bash
# knob sweep: tools/exp_sweep.sh <workload> "ENV1=a ENV2=b" "ENV1=c" ...
w=$1; shift
for cfg in "$@"; do
  echo "### $w [$cfg]"
  env $cfg python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $w 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  us/step', round(d['ms_per_step']*1e3,2), 'p50', round(d['p50_ms']*1e3,2), 'p99', round(d['p99_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3))
"
done
