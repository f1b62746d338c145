import torch, time, os
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
os.system("nvidia-smi topo -m | head -8; lscpu | grep -i -E 'numa|socket|model name' | head -8")
for mb in (11.7, 64):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"H2D {mb} MB: {ms*1e3:.1f} us  {n/ms/1e6:.1f} GB/s")
