#!/usr/bin/env python
"""Where the end-to-end time of the host-buffer entry points goes: device-resident forward time per batch size, the
host-to-device copy alone, and qcnn_net_forward_u8_h / qcnn_net_forward_h for several pipeline chunk sizes."""
import importlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    q = importlib.import_module("quantized-cnn_b200")
    ctx = q.Context(0)
    d, pfx, what = bench.model_files(q, tempfile.mkdtemp())
    net = q.Net(ctx, d, pfx, "AlexNet")
    B = 256
    rs = np.random.RandomState(0)
    pix = torch.from_numpy(rs.randint(0, 256, size=(B, 227, 227, 3)).astype(np.uint8)).pin_memory()
    x32 = torch.from_numpy((rs.rand(B, 3, 227, 227) * 256 - 128).astype(np.float32)).pin_memory()
    net.set_input_mean(np.full((3, 227, 227), 128.0, np.float32))
    xd = x32.cuda()

    def dev_ms(n, reps=20):
        xin = xd[:n].contiguous()
        p = torch.empty((n, 1000), device="cuda")
        for _ in range(3):
            net.forward(xin, prob=p)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            net.forward(xin, prob=p)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    for n in (16, 32, 64, 96, 128, 192, 256):
        ms = dev_ms(n)
        print("device-resident N=%3d: %.3f ms  %.0f img/s" % (n, ms, n / ms * 1e3))
    # copies alone
    for name, t in (("u8 39.6 MB", pix), ("f32 158 MB", x32)):
        dst = torch.empty_like(t, device="cuda")
        for _ in range(3):
            dst.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dst.copy_(t, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 10 * 1e3
        print("H2D %s: %.3f ms  %.1f GB/s" % (name, ms, t.numel() * t.element_size() / ms / 1e6))
    ti = torch.empty((B, 5), dtype=torch.int32).pin_memory()
    tp = torch.empty((B, 5), dtype=torch.float32).pin_memory()
    ph = torch.empty((B, 1000), dtype=torch.float32).pin_memory()
    for chunk in (32, 64, 128, 256):
        net.set_chunk(chunk)
        for name, fn in (("u8+top5", lambda: net.forward_u8_host(pix, k=5, idx_h=ti, val_h=tp)),
                         ("f32+probs", lambda: net.forward_host(x32, ph))):
            for _ in range(4):
                fn()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            ms = (time.perf_counter() - t0) / 10 * 1e3
            print("chunk %3d %-9s: %.3f ms/step  %.0f img/s" % (chunk, name, ms, B / ms * 1e3))
    net.close()
    ctx.close()


if __name__ == "__main__":
    main()
