#!/bin/bash
run() { echo "### $*"; env "$@" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['config']['workload'], 'us/step', round(d['ms_per_step']*1e3,2), 'p50', round(d['p50_ms']*1e3,2), 'frac', round(d['roofline']['frac'],3))
"; }
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline"
run $B --workload model_E
run RECOM_B200_HASH_WEIGHT=5 $B --workload model_E
run RECOM_B200_HASH_WEIGHT=10 $B --workload model_E
run RECOM_B200_HASH_WEIGHT=10 RECOM_B200_POOL_PCT=5 $B --workload model_E
run RECOM_B200_HASH_WEIGHT=10 RECOM_B200_POOL_PCT=25 $B --workload model_E
run RECOM_B200_POOL_PCT=5 $B --workload model_E_bucketize
run RECOM_B200_POOL_PCT=25 $B --workload model_E_bucketize
