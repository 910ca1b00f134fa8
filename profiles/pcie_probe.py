import torch, time
for mb in (1, 4, 16, 64, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    for _ in range(10): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print("%4d MB  D2H %.1f GB/s  H2D %.1f GB/s" % (mb, 10*n/(t1-t0)/1e9, 10*n/(t3-t2)/1e9))
