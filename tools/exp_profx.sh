#!/bin/bash
# level timers of the transform path (make variant NAME=profx DEFS="-DRECOM_PROFILE -DRECOM_PROFILE_X")
for w in "$@"; do
  RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_profx.so RECOM_B200_PROF_DUMP=gpurun_out/profx_$w.txt \
    python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $w > /dev/null 2>gpurun_out/profx_$w.err
  echo "## $w (xform levels: claim+item, slot/rs, offsets, words, hash+store, flag+claim)"
  python - <<PY
import numpy as np
a=np.loadtxt("gpurun_out/profx_$w.txt",dtype=np.uint64).astype(np.float64)
c=a[:,:6]/1965.0
m=c.sum(1)>0
print("warps with tasks",m.sum(),"of",len(a))
for i,n in enumerate(["claim+item","slot/rs","offsets","words","hash+st","flag+claim"]):
    v=c[m,i]; print(f"  {n:<11} mean {v.mean():6.2f} p50 {np.median(v):6.2f} p90 {np.percentile(v,90):6.2f} max {v.max():6.2f} us")
PY
done
