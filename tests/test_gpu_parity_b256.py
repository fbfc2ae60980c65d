"""Parity of the BENCHMARKED configuration: batch 256, where every conv layer runs the decode-at-use tensor-core GEMM with
256-position tiles and ONE TMEM accumulator (pq_gemm_tc NT = 256) and the FC layers run its split-K mode -- a different
numerical path from the small-batch tilings the other tests exercise (two accumulators, K-split).

What is checked, against the CPU oracle (oracle/pq_oracle.c, pinned bit-for-bit to the compiled reference):
  * every conv layer of AlexNet alone at N = 256 with the kernel family PINNED (force_kernel) and the plan asserted from
    qcnn_layer_describe -- so a green run says which kernel was checked;
  * the shipped (else synthetic) AlexNet end to end at N = 256, default and strict: all 24 feature maps of 32 images
    spread over the batch, logits and probabilities of all 256;
  * >= 1024 LCG images: top-1 and top-5 agreement with the reference (stands in for CaffeEva::CalcPredAccu,
    src/CaffeEva.cc:263-295, the dataset being absent).
Errors are reported under BOTH metrics -- e1 = |d| / max(1, |ref|) and e2 = |d| / max(1, |ref|,
0.1 max|ref|) (test_gpu_layers.close) -- into gpurun_out/parity_b256.json; the asserted bounds are stated below."""
import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from test_gpu_layers import RTOL, RTOL_TC, close, rand_act
from test_gpu_net import MODES, set_mode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}
# e1 bound of the default (3xTF32, one accumulator) path: elements of magnitude <= 1 inside maps that reach ~10^3 carry
# the absolute rounding noise of the large sums next to them, so e1 is ~10x e2 (measured values in the report)
E1_TC = 2e-3
E1_STRICT = 2e-4
WORKERS = max(1, min(32, (os.cpu_count() or 2) - 1))


def e1(gpu, ref):
    gpu = np.asarray(gpu, np.float64)
    ref = np.asarray(ref, np.float64)
    return float((np.abs(gpu - ref) / np.maximum(1.0, np.abs(ref))).max())


def save_report():
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_b256.json"), "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def par(fn, items):
    """ctypes calls release the GIL: the oracle runs one image per thread."""
    with ThreadPoolExecutor(WORKERS) as ex:
        return list(ex.map(fn, items))


CONV_LAYERS = {
    # name: (Hi, Wi, Cin, Cout, k, pad, stride, G, S, K, d)
    "conv1": (227, 227, 3, 96, 11, 0, 4, 1, 1, 128, 8),
    "conv2": (27, 27, 96, 256, 5, 2, 1, 2, 6, 128, 8),
    "conv3": (13, 13, 256, 384, 3, 1, 1, 1, 32, 128, 8),
    "conv4": (13, 13, 384, 384, 3, 1, 1, 2, 24, 128, 8),
    "conv5": (13, 13, 384, 256, 3, 1, 1, 2, 24, 128, 8),
}
# kernel families pinned per layer: 6 = pq_gemm_tc (what the bench runs; both accumulator layouts, NT = 256 and 128, are
# pinned in turn) and the strict path's family at the same batch
PINS = {"conv1": [6, 4], "conv2": [6, 2], "conv3": [6, 2], "conv4": [6, 2], "conv5": [6, 2]}
KNAME = {6: "pq_gemm_tc", 4: "conv_direct", 2: "conv_s1_tc", 0: "conv_s1", 1: "conv_roll", 3: "conv_roll_tc"}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CONV_LAYERS))
def test_conv_layer_b256_pinned_kernel(name, po, qcnn, ctx):
    import torch
    Hi, Wi, Cin, Cout, k, pad, stride, G, S, K, d = CONV_LAYERS[name]
    N = 256
    rng = np.random.RandomState(hash(name) % (2 ** 31))
    fan = (Cin // G) * k * k
    ctrd = (rng.randn(S, K, d) / np.sqrt(fan)).astype(np.float32)
    asmt = rng.randint(0, K, size=(Cout, k, k, S)).astype(np.uint8)
    bias = (rng.randn(Cout) * 0.05).astype(np.float32)
    x = rand_act(rng, (N, Hi, Wi, Cin), scale=20.0)
    L = po.conv(pad, k, Cout, G, stride)
    ref = np.concatenate(par(lambda i: po.conv_aprx(x[i:i + 1], L, ctrd, asmt, bias), range(N)))
    layer = qcnn.ConvLayer(ctx, Cin, Hi, Wi, Cout, k, pad, stride, G, ctrd, asmt, bias)
    xd = torch.from_numpy(x).cuda()
    # tensor_core: 1 = 3xTF32 operands, 2 = bf16x2 operands (two MMAs per k-step), 0 = LUT + gather kernels
    for kern, nt, tc in [(6, 256, 1), (6, 128, 1), (6, 256, 2), (6, 128, 2)] + [(k_, 0, 0) for k_ in PINS[name] if k_ != 6]:
        layer.set_param("tensor_core", tc)
        layer.set_param("force_kernel", kern)
        layer.set_param("gemm_nt", nt)
        y = layer.forward(xd).cpu().numpy()
        desc = layer.describe(N)          # after the forward: the plan the autotuner settled on and just ran
        assert KNAME[kern] in desc, desc
        if kern == 6:
            # NT=256: ONE TMEM accumulator (all 3xTF32 terms chained in it); NT=128: the cross terms have their own
            assert "NT=%d" % nt in desc and "nsplit=1" in desc, desc
            assert ("bf16x2" in desc) == (tc == 2), desc
        a, b = e1(y, ref), close(y, ref)
        REPORT["%s/N=256/%s%s%s" % (name, KNAME[kern], "/NT=%d" % nt if nt else "", "/bf16x2" if tc == 2 else "")] = {
            "e1_max1ref": a, "e2_close": b, "plan": desc}
        assert b <= (RTOL_TC if kern == 6 else RTOL), (name, kern, nt, b)
        assert a <= (E1_TC if kern == 6 else E1_STRICT), (name, kern, nt, a)
        yr = layer.forward(xd, relu=True).cpu().numpy()
        assert close(yr, np.maximum(ref, 0)) <= (RTOL_TC if kern == 6 else RTOL)
    # without a pin the default family at this batch is the tensor-core GEMM (family = setting, not a timing result)
    layer.set_param("force_kernel", -1)
    layer.set_param("gemm_nt", 0)
    layer.set_param("tensor_core", 1)
    assert "pq_gemm_tc" in layer.describe(N)
    layer.set_param("tensor_core", 0)
    assert "pq_gemm_tc" not in layer.describe(N)
    layer.close()
    save_report()


@pytest.fixture(scope="module")
def alexnet_b256(po, tmp_path_factory):
    """Model files + 256 LCG images + oracle feature maps of 32 of them and logits / probabilities of all."""
    if po.have_alexnet():
        dirpath, pfx, what = po.ALEXNET_DIR, po.ALEXNET_PFX, "shipped"
    else:
        dirpath, pfx, what = str(tmp_path_factory.mktemp("synth_b256")), "synth", "synthetic"
        po.save_model(dirpath, pfx, po.synth_alexnet(seed=1))
    layers = po.alexnet_layers()
    params = po.load_model(dirpath, pfx, layers)
    N = 256
    img = po.lcg_images(N, 777000)
    keep = list(range(0, N, 8))      # 32 images, every 8th: their positions fall on every alignment of the 256-position tiles

    def one(i):
        prob, maps = po.net_forward(layers, params, img[i:i + 1], keep=True)
        return prob[0], maps[22].reshape(-1), ([m[0] for m in maps] if i in keep else None)
    res = par(one, range(N))
    return dict(dirpath=dirpath, pfx=pfx, what=what, img=img, keep=keep, prob=np.stack([r[0] for r in res]),
                logits=np.stack([r[1] for r in res]), maps={i: res[i][2] for i in keep})


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["default", "strict"])
def test_alexnet_b256_all_feature_maps(mode, po, qcnn, ctx, alexnet_b256):
    import torch
    g = alexnet_b256
    N = 256
    net = qcnn.Net(ctx, g["dirpath"], g["pfx"], "AlexNet")
    set_mode(net, mode)
    RT, PT = MODES[mode]
    imgd = torch.from_numpy(g["img"]).cuda()
    logits = torch.empty((N, 1000), dtype=torch.float32, device="cuda")
    net.forward(imgd)      # tilings of this batch size are timed here; describe() then reports the plans that run
    plans = {}
    for l in (0, 4, 8, 10, 12, 15, 18, 21):
        desc = net.pq_layer(l).describe(N)
        plans[str(l)] = desc.split(" smem")[0]
        # the family is a setting: tensor-core GEMMs in default mode, LUT + gather kernels in strict mode
        assert ("pq_gemm_tc" in desc) == (mode == "default"), (l, desc)
    # (a) un-fused: all 24 feature maps of the 32 kept images
    net.set_keep_maps(True)
    prob_d = net.forward(imgd, logits=logits)
    worst = {}
    for l in range(24):
        fm = net.featmap(l, N).cpu().numpy()
        ref = np.stack([g["maps"][i][l].reshape(fm.shape[1:]) for i in g["keep"]])
        got = fm[g["keep"]]
        worst[l] = (e1(got, ref), close(got, ref))
        assert worst[l][1] <= RT, (mode, l, worst[l])
    assert close(logits.cpu().numpy(), g["logits"]) <= RT
    assert np.abs(prob_d.cpu().numpy() - g["prob"]).max() <= PT
    # (b) the fused production path (what bench.py times)
    net.set_keep_maps(False)
    prob = net.forward(imgd, logits=logits).cpu().numpy()
    lg = logits.cpu().numpy()
    assert close(lg, g["logits"]) <= RT
    assert np.abs(prob - g["prob"]).max() <= PT
    REPORT["alexnet(%s)/N=256/%s" % (g["what"], mode)] = {
        "feature_maps_e1_max": max(v[0] for v in worst.values()), "feature_maps_e2_max": max(v[1] for v in worst.values()),
        "per_map_e2": {str(l): worst[l][1] for l in worst}, "logits_e1": e1(lg, g["logits"]),
        "logits_e2": close(lg, g["logits"]), "prob_abs": float(np.abs(prob - g["prob"]).max()),
        "top1_agree": int((prob.argmax(1) == g["prob"].argmax(1)).sum()), "images": N, "plans": plans}
    net.close()
    save_report()


def _topk_sets(p, k=5):
    idx = np.argsort(-p, axis=1, kind="stable")[:, :k + 1]
    return idx


@pytest.mark.gpu
def test_top5_agreement_on_1024_images(po, qcnn, ctx, alexnet_b256):
    """Stand-in for the 1k-image accuracy run (the dataset tensor is absent): identical top-1 and top-5 SET on 1024 seeded
    inputs, both paths; a disagreement is accepted only where the reference's own probabilities are tied to within the
    path's tolerance (then either order is a correct answer of the fp32 computation)."""
    import torch
    g = alexnet_b256
    layers = po.alexnet_layers()
    params = po.load_model(g["dirpath"], g["pfx"], layers)
    M = 1024
    img = po.lcg_images(M, 31337)
    ref = np.stack(par(lambda i: po.net_forward(layers, params, img[i:i + 1])[0], range(M)))
    net = qcnn.Net(ctx, g["dirpath"], g["pfx"], "AlexNet")
    rep = {}
    for mode in ("default", "strict"):
        set_mode(net, mode) if mode == "strict" else None
        PT = MODES[mode][1]
        prob = np.concatenate([net.forward(torch.from_numpy(img[b:b + 256]).cuda()).cpu().numpy() for b in range(0, M, 256)])
        ir, ig = _topk_sets(ref), _topk_sets(prob)
        top1_same = ir[:, 0] == ig[:, 0]
        top5_same = np.array([set(ir[i, :5]) == set(ig[i, :5]) for i in range(M)])
        srt = -np.sort(-ref, axis=1)
        tie1 = (srt[:, 0] - srt[:, 1]) <= 4 * PT
        tie5 = (srt[:, 4] - srt[:, 5]) <= 4 * PT
        assert (top1_same | tie1).all(), (mode, np.where(~(top1_same | tie1))[0][:8])
        assert (top5_same | tie5).all(), (mode, np.where(~(top5_same | tie5))[0][:8])
        rep[mode] = {"images": M, "top1_identical": int(top1_same.sum()), "top5_set_identical": int(top5_same.sum()),
                     "top1_ties_in_reference": int(tie1.sum()), "top5_ties_in_reference": int(tie5.sum()),
                     "prob_abs_max": float(np.abs(prob - ref).max())}
        assert np.abs(prob - ref).max() <= PT
    REPORT["top5_agreement(%s)" % g["what"]] = rep
    net.close()
    save_report()
