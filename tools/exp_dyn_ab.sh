#!/bin/bash
# config 3 (dynamic): long-row kernel variants - register budget (resident CTAs) x CTAs per SM
mkdir -p gpurun_out/r2
run() { env $1 python bench.py --workload dynamic --batch $2 --steps 100 --warmup 5 --no-cpu-baseline 2>>gpurun_out/r2/dyn4.err | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', d['config']['workload'], round(d['ms_per_step']*1e3,1), round(d['p99_ms']*1e3,1))"; }
for lib in librecom_b200 librecom_b200_long3 librecom_b200_long4; do
  for c in 4 6 8; do run "RECOM_B200_LIB=$PWD/recom_b200/$lib.so RECOM_B200_LONG_CTAS=$c" 2048; done
done
run "RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_long4.so RECOM_B200_LONG_CTAS=8" 8192
run "RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_long3.so RECOM_B200_LONG_CTAS=6" 8192
