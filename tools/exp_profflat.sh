#!/bin/bash
# level timers of the flat kernel's column tiles (profc variant, RECOM_B200_PROF_FLAT=1)
for w in "$@"; do
  RECOM_B200_PROF_FLAT=1 RECOM_B200_LIB=$PWD/recom_b200/librecom_b200_profc.so RECOM_B200_PROF_DUMP=gpurun_out/proff_$w.txt \
    python bench.py --steps 50 --warmup 5 --no-cpu-baseline --workload $w > /dev/null 2>gpurun_out/proff_$w.err
  echo "## $w (flat kernel, column tile levels, us per warp)"
  python - <<PY
import numpy as np
a=np.loadtxt("gpurun_out/proff_$w.txt",dtype=np.uint64).astype(np.float64)
a=a[a[:,6]>0]
c=a[:,:6]/1965.0
t0=a[:,6].min()
print("tiles",len(a)," span %.1f us"%((a[:,7].max()-t0)/1e3), " start p50 %.1f p90 %.1f max %.1f"%tuple(np.percentile((a[:,6]-t0)/1e3,[50,90,100])),
      " dur p50 %.1f p90 %.1f max %.1f"%tuple(np.percentile((a[:,7]-a[:,6])/1e3,[50,90,100])))
for i,n in enumerate(["item","resolve","row_splits","transform","gather"]):
    v=c[:,i]; print(f"  {n:<11} mean {v.mean():6.2f} p50 {np.median(v):6.2f} p90 {np.percentile(v,90):6.2f} max {v.max():6.2f}")
# the first 1/4 of the tiles (most expensive) separately
q=len(a)//4
for nm,sl in (("first quarter",slice(0,q)),("last quarter",slice(len(a)-q,len(a)))):
    v=c[sl]; d=(a[sl,7]-a[sl,6])/1e3
    print(f"  {nm}: dur mean {d.mean():.1f}; levels "+" ".join(f"{v[:,i].mean():.2f}" for i in range(5)))
PY
done
