"""Chained (whole-network) error of the device executor against the CPU oracle, per feature map, for the default
configuration (tensor-core kernels allowed) and the strict one (tensor_core = 0).   python tools/net_accuracy.py [N]"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def close(gpu, ref):
    gpu = np.asarray(gpu, np.float64)
    ref = np.asarray(ref, np.float64)
    scale = np.maximum(np.maximum(1.0, np.abs(ref)), 0.1 * np.abs(ref).max())
    return float((np.abs(gpu - ref) / scale).max())


def main():
    import torch
    import pyoracle as po
    q = importlib.import_module("quantized-cnn_b200")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    d = tempfile.mkdtemp(prefix="qcnn_acc_")
    params = po.synth_alexnet(seed=1)
    po.save_model(d, "synth", params)
    layers = po.alexnet_layers()
    img = po.lcg_images(N, 4242)
    ref_prob, ref_maps = po.net_forward(layers, params, img, keep=True)
    ctx = q.Context(0)
    for mode in ("default", "strict"):
        net = q.Net(ctx, d, "synth", "AlexNet")
        if mode == "strict":
            for l in range(net.layer_count):
                pl = net.pq_layer(l)
                if pl is not None:
                    pl.set_param("tensor_core", 0)
        net.set_keep_maps(True)
        imgd = torch.from_numpy(img).cuda()
        logits = torch.empty((N, 1000), dtype=torch.float32, device="cuda")
        prob_d = net.forward(imgd, logits=logits)
        prob_d = net.forward(imgd, logits=logits)
        errs = []
        for l in range(24):
            fm = net.featmap(l, N)
            errs.append(close(fm.cpu().numpy().reshape(-1), ref_maps[l].reshape(-1)))
        print(mode, "N=%d" % N, "maps max %.3g at %d |" % (max(errs), int(np.argmax(errs))), " ".join("%.2g" % e for e in errs))
        print(mode, "prob abs err %.3g, logits metric %.3g" % (np.abs(prob_d.cpu().numpy() - ref_prob).max(), close(logits.cpu().numpy(), ref_maps[22])))
        for l in range(net.layer_count):
            pl = net.pq_layer(l)
            if pl is not None and mode == "default":
                print("   layer", l, pl.describe(N)[:60])
        net.close()


if __name__ == "__main__":
    main()
