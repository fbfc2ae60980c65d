"""Pins the CPU oracle (oracle/pq_oracle.c): bit-for-bit against
  (1) the committed golden vectors generated from the compiled reference (tests/golden/make_golden.py), always;
  (2) the compiled reference itself (oracle/_ref/libqcnn_ref.so) when it is present (it is built from the
      unmodified /root/reference sources by oracle/Makefile and travels git-ignored).
No GPU, no product code."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLD, name))


SYNTH = {
    # must mirror tests/golden/make_golden.py
    "convA": ("conv", (1, 3, 32, 1, 1), (16, 7, 9)),
    "convB": ("conv", (2, 5, 32, 2, 1), (12, 9, 9)),
    "convC": ("conv", (0, 7, 16, 1, 3), (3, 23, 23)),
    "fcA": ("fc", 40, (24, 1, 1)),
    "fcB": ("fc", 24, (10, 1, 1)),
}


@pytest.mark.parametrize("name", sorted(SYNTH))
def test_oracle_single_layer_matches_reference_golden(name, po):
    g = golden("synth_layers.npz")
    kind, spec, chw = SYNTH[name]
    img = g[name + "_img"][None]
    ctrd, asmt, bias = g[name + "_ctrd0"], g[name + "_asmt0"], g[name + "_bias0"]
    x = po.nchw_to_nhwc(img)
    assert np.array_equal(x, g[name + "_fm0"])
    if kind == "conv":
        y = po.conv_aprx(x, po.conv(*spec), ctrd, asmt, bias)
    else:
        y = po.fc_aprx(po.nhwc_to_nchw(x).reshape(1, -1), ctrd, asmt, bias)
    assert np.array_equal(y.reshape(-1), g[name + "_fm1"].reshape(-1))


def test_oracle_misc_layers_match_reference_golden(po):
    g = golden("synth_layers.npz")
    x = g["misc_fm0"]
    r = po.relu_f(x)
    assert np.array_equal(r, g["misc_fm1"])
    n = po.lrn_f(r, 5, 1e-4, 0.75, 1.0)
    assert np.array_equal(n, g["misc_fm2"])
    p = po.pool_f(n, 3, 0, 2)
    assert np.array_equal(p, g["misc_fm3"])


def test_oracle_tiny_net_matches_reference_golden(po):
    g = golden("synth_layers.npz")
    layers = [po.conv(1, 3, 16, 1, 1), po.relu(), po.pool(0, 2, 2), po.fcnt(16), po.relu(), po.drpt(0.5),
              po.fcnt(8), po.smax()]
    params = {l: dict(bias=g["tiny_bias%d" % l], ctrd=g["tiny_ctrd%d" % l], asmt=g["tiny_asmt%d" % l]) for l in (0, 3, 6)}
    prob, maps = po.net_forward(layers, params, g["tiny_img"][None], keep=True)
    assert np.array_equal(prob.reshape(-1), g["tiny_out"])
    for l in (0, 1, 2, 4, 5, 6, 7, 8):  # featMapLst[3] is left in NCHW element order by the reference
        assert np.array_equal(maps[l].reshape(-1), g["tiny_fm%d" % l].reshape(-1)), l
    assert np.array_equal(po.nhwc_to_nchw(maps[3]).reshape(-1), g["tiny_fm3"].reshape(-1))
    assert abs(float(prob.sum()) - 1.0) < 1e-5


def test_oracle_file_formats_match_reference_bytes(po, tmp_path):
    g = golden("cbn_vectors.npz")
    keys = sorted(k[:-5] for k in g.files if k.endswith("_idx0"))
    assert len(keys) == 6
    for key in keys:
        idx0, blob = g[key + "_idx0"], g[key + "_file"]
        bits = int(key[1])
        ref_path = str(tmp_path / (key + ".ref.cbn"))
        blob.tofile(ref_path)
        got, b = po.read_cbn(ref_path)
        assert b == bits and np.array_equal(got, idx0)
        mine = str(tmp_path / (key + ".mine.cbn"))
        po.write_cbn(mine, idx0, bits)
        assert np.array_equal(np.fromfile(mine, np.uint8), blob)
    p = str(tmp_path / "t.bin")
    g["bin_file"].tofile(p)
    assert np.array_equal(po.read_bin(p), g["bin_arr"])
    po.write_bin(p + "2", g["bin_arr"])
    assert np.array_equal(np.fromfile(p + "2", np.uint8), g["bin_file"])


def test_lcg_image_generator(po):
    img = po.lcg_images(1, 12345).reshape(-1)
    s = 12345
    exp = []
    for _ in range(8):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        exp.append(((s >> 8) & 0xFFFF) / 65536.0 * 256.0 - 128.0)
    assert np.array_equal(img[:8], np.array(exp, np.float32))
    assert img.min() >= -128 and img.max() < 128


def test_oracle_alexnet_kat_matches_reference_golden(po):
    """Shipped quantized AlexNet (staged under oracle/_ref/data by oracle/Makefile) through the oracle port."""
    if not po.have_alexnet():
        pytest.skip("shipped AlexNet parameters not staged (oracle/_ref/data)")
    g = golden("alexnet_kat.npz")
    layers = po.alexnet_layers()
    params = po.load_model(po.ALEXNET_DIR, po.ALEXNET_PFX, layers)
    imgs = po.lcg_images(2, 12345)
    for i in range(2):
        prob, maps = po.net_forward(layers, params, imgs[i:i + 1], keep=True)
        assert np.array_equal(prob[0], g["prob%d" % i])
        assert np.array_equal(maps[22].reshape(-1), g["logits%d" % i])
        idx, _ = po.topk(prob[0], 5)
        assert np.array_equal(idx, g["top5_%d" % i])
        for l in (1, 5, 9, 11, 13, 16, 19, 22):
            assert np.array_equal(maps[l].reshape(-1)[:64], g["fm%d_%d" % (l, i)])
        cks = g["cks%d" % i]
        for l in range(24):
            m = maps[l].astype(np.float64).reshape(-1)
            assert np.allclose([m.sum(), np.sqrt((m * m).sum()), m.max()], cks[l], rtol=1e-12, atol=0)
    # seed 12345 -> class 533, p = 0.621259
    assert int(g["top5_0"][0]) == 533 and abs(float(g["prob0"][533]) - 0.621259) < 1e-6


# ---- live cross-checks against the compiled reference (present in the build container and on the GPU box) ----
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref",
                                                               "libqcnn_ref.so")), reason="oracle/_ref not built")


@needs_ref
def test_oracle_lut_stage_matches_live_reference(po):
    rng = np.random.RandomState(11)
    for (P, D, S, K, d) in [(50, 48, 6, 128, 8), (7, 3, 1, 128, 8), (3, 10, 3, 32, 4), (2, 4096, 4096, 16, 1)]:
        data = (rng.randn(P, D) * 10).astype(np.float32)
        ctrd = (rng.randn(S, K, d) * 0.1).astype(np.float32)
        assert np.array_equal(po.get_inpd(data, ctrd), po.ref_get_inpd(data, ctrd))


@needs_ref
def test_oracle_random_layers_match_live_reference(po, tmp_path):
    """Seeded random conv / FC shapes through the reference's own CalcFeatMap_ConvAprx / _FCntAprx."""
    rng = np.random.RandomState(99)
    cases = [
        ([po.conv(1, 3, 64, 2, 1)], (32, 13, 13), {0: (4, 64, 4)}),
        ([po.conv(2, 5, 48, 1, 2)], (6, 17, 15), {0: (2, 128, 4)}),      # d > remaining dims in last subspace
        ([po.conv(0, 11, 32, 1, 4)], (3, 51, 51), {0: (1, 128, 8)}),
        ([po.fcnt(64)], (30, 2, 2), {0: (30, 32, 4)}),
        ([po.fcnt(1000)], (100, 1, 1), {0: (100, 16, 1)}),
    ]
    for ci, (layers, chw, pq) in enumerate(cases):
        params = po.synth_model(layers, chw, pq, seed=ci, ctrd_std=0.2)
        d = str(tmp_path / ("m%d" % ci))
        po.save_model(d, "rnd", params)
        net = po.RefNet(d, "rnd", layers=layers, in_chw=chw)
        # decoded parameters identical (bit-exact assignment indexing)
        a, _ = net.param(0, 2)
        assert np.array_equal(a, params[0]["asmt"].reshape(-1))
        for _ in range(2):
            img = (rng.randn(*chw) * 5).astype(np.float32)
            ref = net.forward(img)
            mine = po.net_forward(layers, params, img[None])
            assert np.array_equal(mine.reshape(-1), ref)
        net.close()


@needs_ref
def test_alexnet_live_reference_vs_oracle(po):
    if not po.have_alexnet():
        pytest.skip("shipped AlexNet parameters not staged")
    net = po.RefNet(po.ALEXNET_DIR, po.ALEXNET_PFX)
    layers = po.alexnet_layers()
    params = po.load_model(po.ALEXNET_DIR, po.ALEXNET_PFX, layers)
    img = po.lcg_images(1, 777)
    ref = net.forward(img[0])
    mine, maps = po.net_forward(layers, params, img, keep=True)
    assert np.array_equal(mine[0], ref)
    for l in range(24):
        if l == 15:
            continue  # reference leaves featMapLst[15] in NCHW order (CaffeEva.cc:246-253)
        assert np.array_equal(maps[l].reshape(-1), net.featmap(l).reshape(-1)), l
    for l in params:
        for which, key in ((0, "bias"), (1, "ctrd"), (2, "asmt")):
            a, _ = net.param(l, which)
            assert np.array_equal(a, params[l][key].reshape(-1))
    net.close()
