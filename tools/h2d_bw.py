import torch, time
for mb in (39, 158, 632):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("H2D pinned %d MB: %.1f GB/s" % (mb, 5 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
    e0.record()
    for _ in range(5): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print("D2H pinned %d MB: %.1f GB/s" % (mb, 5 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
