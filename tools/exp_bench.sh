#!/bin/bash
# experiment runner: several bench configurations in one gpurun call
run() { echo "### $*"; env "$@" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config']['workload'], 'us/step', round(d['ms_per_step']*1e3,2), 'p50', round(d['p50_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'e2e_us', round(d['e2e']['ms_per_step']*1e3,1))
"; }
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline"
for w in "$@"; do run $B --workload $w; done
