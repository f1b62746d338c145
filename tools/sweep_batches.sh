#!/bin/bash
# batch sweeps with p50/p99, achieved GB/s and the CPU baseline beside each
# line (bench.py's own legs: device-timed, e2e, oracle parity of a timed request, CPU oracle port).
#   tools/sweep_batches.sh > gpurun_out/r2/sweep.jsonl
run() {  # workload batch steps cpu_seconds
  python bench.py --workload $1 --batch $2 --steps $3 --warmup 5 --cpu-seconds $4 2>/dev/null | tail -1
}
for b in 256 512 1024 2048 4096; do run model_E $b 200 3; done
for b in 256 512 1024 2048 4096; do run model_F $b 200 3; done
for b in 1 16 128 1024 2048 4096 8192; do run dynamic $b 200 2; done
for w in mixed1000 mixed1000_bf16 production2000 model_E_bf16 micro100; do run $w 0 200 3; done
