import torch, time
for mb in (1, 16, 256):
    a = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    b = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    n = max(4, 2048 // mb)
    t0 = time.perf_counter()
    for _ in range(n): b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {mb} MiB pinned: {n*mb*1.048576e-3/dt:.1f} GB/s", flush=True)
