"""Pinned host -> device copy bandwidth: one stream vs several concurrent streams (what bounds bench.py's e2e number)."""
import torch

def run(total_mb, nstreams, reps=5):
    n = total_mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    part = n // nstreams
    def go():
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * part:(i + 1) * part].copy_(h[i * part:(i + 1) * part], non_blocking=True)
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        go()
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return reps * part * nstreams * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9

for mb in (39, 158, 632):
    print("H2D pinned %4d MB:" % mb, " ".join("%d streams %.1f GB/s;" % (k, run(mb, k)) for k in (1, 2, 4, 8)))

# NUMA placement of the pinned buffer: bind the allocating thread to the CPUs next to the GPU (NVML's ideal affinity)
try:
    import os
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
    print("cpus allowed before:", len(os.sched_getaffinity(0)))
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("cpus allowed after nvmlDeviceSetCpuAffinity:", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
    for mb in (39, 158, 158, 158, 632):
        print("bound: H2D pinned %4d MB:" % mb, " ".join("%d streams %.1f GB/s;" % (k, run(mb, k)) for k in (1, 2)))
    pynvml.nvmlDeviceClearCpuAffinity(h)
    for mb in (158, 158, 158):
        print("unbound again: H2D pinned %4d MB:" % mb, " ".join("%d streams %.1f GB/s;" % (k, run(mb, k)) for k in (1, 2)))
except Exception as e:  # noqa: BLE001
    print("affinity test skipped:", e)
