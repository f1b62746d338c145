#!/bin/bash
# compact bench lines: tools/exp_run.sh <label> "<ENV=.. ENV=..>" workload...   -> gpurun_out/r2/<label>.jsonl
label=$1; envs=$2; shift 2
mkdir -p gpurun_out/r2; out=gpurun_out/r2/$label.jsonl; : > $out
echo "### $label [$envs]"
for w in "$@"; do
  env $envs python bench.py --steps 100 --warmup 5 --no-cpu-baseline --workload $w 2>>gpurun_out/r2/err.log | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps({'w':d['config']['workload'],'us':round(d['ms_per_step']*1e3,2),'p50':round(d['p50_ms']*1e3,2),'frac':round(d['roofline']['frac'],3),'e2e_us':round(d['e2e']['ms_per_step']*1e3,1)}))
" | tee -a $out
done
