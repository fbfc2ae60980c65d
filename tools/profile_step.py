"""Minimal driver for ncu captures: builds the AlexNet PQ net and runs `--iters` forward passes at batch `--batch`.
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_|fc_aprx|lrn_' -o out \
        python tools/profile_step.py --batch 256 --iters 2
"""
import argparse
import importlib
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--tc", type=int, default=-1, help="tensor_core parameter for every PQ layer")
    args = ap.parse_args()
    import torch
    q = importlib.import_module("quantized-cnn_b200")
    tmp = tempfile.mkdtemp(prefix="qcnn_prof_")
    d, pfx, what = bench.model_files(q, tmp)
    ctx = q.Context(0)
    net = q.Net(ctx, d, pfx, "AlexNet")
    if args.tc >= 0:
        for l in range(net.layer_count):
            if net.pq_layer(l) is not None:
                net.pq_layer(l).set_param("tensor_core", args.tc)
    img = torch.from_numpy(bench.lcg_images(args.batch, 12345)).cuda()
    prob = torch.empty((args.batch, 1000), dtype=torch.float32, device="cuda")
    for l in range(net.layer_count):
        pl = net.pq_layer(l)
        if pl is not None:
            print("layer", l, pl.describe(args.batch))
    # warm-up passes (also lets the conv tilings be autotuned); only the LAST pass is inside the profiler range:
    #   ncu --profile-from-start off ... python tools/profile_step.py
    for _ in range(max(args.iters - 1, 1)):
        net.forward(img, prob=prob)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.forward(img, prob=prob)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("done", what, float(prob[0].max()))


if __name__ == "__main__":
    main()
