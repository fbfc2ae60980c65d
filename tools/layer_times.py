"""Per-layer CUDA-event times + chosen tilings of the AlexNet PQ net at a given batch size.
    python tools/layer_times.py --batch 1 [--reps 50]"""
import argparse
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--tc", type=int, default=-1, help="tensor_core parameter for every PQ layer (0 strict, 1 3xTF32, 2 bf16x2)")
    args = ap.parse_args()
    import torch
    q = importlib.import_module("quantized-cnn_b200")
    d, pfx, what = bench.model_files(q, tempfile.mkdtemp(prefix="qcnn_lt_"))
    ctx = q.Context(0)
    net = q.Net(ctx, d, pfx, "AlexNet")
    B = args.batch
    if args.tc >= 0:
        for l in range(net.layer_count):
            if net.pq_layer(l) is not None:
                net.pq_layer(l).set_param("tensor_core", args.tc)
    img = torch.from_numpy(bench.lcg_images(B, 12345)).cuda()
    prob = torch.empty((B, 1000), dtype=torch.float32, device="cuda")
    for _ in range(5):
        net.forward(img, prob=prob)
    torch.cuda.synchronize()
    # whole-pass time without per-layer events
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        net.forward(img, prob=prob)
    e1.record()
    torch.cuda.synchronize()
    print("batch %d: %.4f ms / forward (%d launches)" % (B, e0.elapsed_time(e1) / args.reps, net.launch_count()))
    net.set_profiling(True)
    acc = np.zeros(net.layer_count)
    for _ in range(args.reps):
        net.forward(img, prob=prob)
        torch.cuda.synchronize()
        acc += np.array([net.layer_time_ms(l) for l in range(net.layer_count)])
    acc /= args.reps
    for l in range(net.layer_count):
        if acc[l] > 0:
            pl = net.pq_layer(l)
            print("  layer %2d  %8.4f ms  %s" % (l, acc[l], pl.describe(B) if pl is not None else ""))
    print("  sum of layers %.4f ms" % acc.sum())


if __name__ == "__main__":
    main()
