#!/bin/bash
# level timers of the column tiles (make variant NAME=profc DEFS="-DRECOM_PROFILE -DRECOM_PROFILE_COL")
for w in "$@"; do
  RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_profc.so RECOM_B200_PROF_DUMP=gpurun_out/profc_$w.txt \
    python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $w > /dev/null 2>gpurun_out/profc_$w.err
  echo "## $w (column tile levels, us per warp)"
  python - <<PY
import numpy as np
a=np.loadtxt("gpurun_out/profc_$w.txt",dtype=np.uint64).astype(np.float64)
c=a[:,:6]/1965.0
m=c[:,3]>0
print("warps with column tiles",m.sum(),"of",len(a), " span %.1f us"%((a[:,7].max()-a[:,6].min())/1e3))
for i,n in enumerate(["claim+item","resolve","row_splits","transform","gather"]):
    v=c[m,i]; print(f"  {n:<11} mean {v.mean():6.2f} p50 {np.median(v):6.2f} p90 {np.percentile(v,90):6.2f} max {v.max():6.2f}")
PY
done
