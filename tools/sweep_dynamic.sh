for b in 16 128 1024 2048 4096 8192; do python bench.py --workload dynamic --batch $b --steps 200 --warmup 5 --cpu-seconds 2 2>gpurun_out/r2/dyn_$b.err | tail -1; done
