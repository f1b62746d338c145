#!/bin/bash
# The measurements of the final tree of the round, in one gpurun call: tools/final_measure.sh
# -> gpurun_out/r2f/ (copied to profiles/r2_* by hand)
O=gpurun_out/r2f; mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu.txt
python bench.py --steps 300 --warmup 10 > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err
: > $O/workloads.jsonl
for w in model_E model_F model_E_bucketize model_E_dense model_E_sparse mixed1000 mixed1000_bf16 production2000 model_E_bf16 dynamic micro100; do
  python bench.py --steps 200 --warmup 10 --no-cpu-baseline --workload $w 2>>$O/workloads.err | tail -1 >> $O/workloads.jsonl
done
# launch lists of the timed region (bench.py brackets it with cudaProfilerStart/Stop)
for w in model_E mixed1000 production2000 dynamic; do
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$w.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --workload $w > /dev/null 2>>$O/ncu.err
done
ncu --set full --clock-control none --import-source on -k regex:fused_columns --launch-skip 4 -c 1 -o $O/final_model_E \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --workload model_E > /dev/null 2>>$O/ncu.err
ncu --set full --clock-control none --import-source on -k regex:column_tiles --launch-skip 4 -c 1 -o $O/final_mixed1000 \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-parity --workload mixed1000 > /dev/null 2>>$O/ncu.err
ncu --set full --clock-control none --import-source on -k regex:dense_relu --launch-skip 8 -c 1 -o $O/final_dense \
  python tools/bench_dense.py > $O/dense_under_ncu.json 2>>$O/ncu.err
python tools/bench_dense.py > $O/dense_2048.json; python tools/bench_dense.py 8192 8120 1024 > $O/dense_8192.json
tail -2 $O/pytest_gpu.txt; cut -c1-300 $O/bench_n1.json; tail -2 $O/ncu.err
