#!/bin/bash
# phase-timer run with a given library and env: tools/exp_prof2.sh <lib.so> "<ENV..>" workload...
lib=$1; envs=$2; shift 2
for w in "$@"; do
  env $envs RECOM_B200_LIB=$PWD/recom_b200/$lib RECOM_B200_PROF_DUMP=gpurun_out/prof_$w.txt \
    python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $w > /dev/null 2>gpurun_out/prof_$w.err
  echo "## $w [$lib $envs]"; python tools/prof_summary.py gpurun_out/prof_$w.txt
done
