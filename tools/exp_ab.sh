#!/bin/bash
# A/B of library variants inside one gpurun call: tools/exp_ab.sh "<workloads>" <variant names...> ("" = default lib)
wl=$1; shift
for rep in 1 2; do
for v in "$@"; do
  lib=$PWD/recom_b200/librecom_b200${v:+_$v}.so
  for w in $wl; do
    RECOM_B200_LIB=$lib python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $w 2>>gpurun_out/ab.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('rep$rep %-10s %-20s us/step %7.2f p50 %7.2f' % ('${v:-default}', d['config']['workload'], d['ms_per_step']*1e3, d['p50_ms']*1e3))
"
  done
done
done
