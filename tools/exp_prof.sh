#!/bin/bash
# phase-timer runs (make prof build) for the given workloads
for w in "$@"; do
  RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_prof.so RECOM_B200_PROF_DUMP=gpurun_out/prof_$w.txt \
    python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $w > /dev/null 2>gpurun_out/prof_$w.err
  echo "## $w"; python tools/prof_summary.py gpurun_out/prof_$w.txt
done
