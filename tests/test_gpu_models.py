"""The other layer tables of CaffePara (reference src/CaffePara.cc:54-237) end to end on the GPU: CaffeNet (pool before
LRN: the two are not fused) and VggCnnS (224x224 input, 7x7 stride-2 first conv, pooling 3/3 and 2/2, 18432-wide fc6),
built BY NAME through qcnn_net_create -- so the product's own tables are what is checked -- with synthetic PQ parameters
in the reference's file formats, against the CPU oracle driven by an independent transcription of the reference tables.
No shipped parameters exist for these models; random-init codebooks / assignments stand in."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from test_gpu_layers import close
from test_gpu_net import MODES, set_mode


def caffenet(po):
    return [po.conv(0, 11, 96, 1, 4), po.relu(), po.pool(0, 3, 2), po.lorn(5, 1e-4, 0.75, 1.0),
            po.conv(2, 5, 256, 2, 1), po.relu(), po.pool(0, 3, 2), po.lorn(5, 1e-4, 0.75, 1.0),
            po.conv(1, 3, 384, 1, 1), po.relu(), po.conv(1, 3, 384, 2, 1), po.relu(), po.conv(1, 3, 256, 2, 1), po.relu(),
            po.pool(0, 3, 2), po.fcnt(4096), po.relu(), po.drpt(0.5), po.fcnt(4096), po.relu(), po.drpt(0.5), po.fcnt(1000),
            po.smax()]


def vggcnns(po):
    return [po.conv(0, 7, 96, 1, 2), po.relu(), po.lorn(5, 5e-4, 0.75, 2.0), po.pool(0, 3, 3),
            po.conv(1, 5, 256, 1, 1), po.relu(), po.pool(0, 2, 2),
            po.conv(1, 3, 512, 1, 1), po.relu(), po.conv(1, 3, 512, 1, 1), po.relu(), po.conv(1, 3, 512, 1, 1), po.relu(),
            po.pool(0, 3, 3), po.fcnt(4096), po.relu(), po.drpt(0.5), po.fcnt(4096), po.relu(), po.drpt(0.5), po.fcnt(1000),
            po.smax()]


MODELS = {
    # name: (layer table, input CHW, {layerInd: (S, K, d)})
    "CaffeNet": (caffenet, (3, 227, 227), {0: (1, 128, 8), 4: (6, 128, 8), 8: (32, 128, 8), 10: (24, 128, 8), 12: (24, 128, 8),
                                           15: (2304, 32, 4), 18: (1024, 32, 4), 21: (4096, 16, 1)}),
    "VggCnnS": (vggcnns, (3, 224, 224), {0: (1, 128, 4), 4: (12, 128, 8), 7: (32, 128, 8), 9: (64, 128, 8), 11: (64, 128, 8),
                                         14: (4608, 32, 4), 17: (1024, 32, 4), 20: (4096, 16, 1)}),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(MODELS))
def test_model_table_end_to_end(name, po, qcnn, ctx, tmp_path):
    import torch
    table, chw, pq = MODELS[name]
    layers = table(po)
    params = po.synth_model(layers, chw, pq, seed=11, ctrd_std=None, bias_std=0.05)
    last = max(pq)
    params[last]["ctrd"] = (params[last]["ctrd"] * np.float32(0.25)).astype(np.float32)   # logits inside expf's range
    d = str(tmp_path)
    po.save_model(d, "m", params)
    net = qcnn.Net(ctx, d, "m", name)
    assert net.layer_count == len(layers) and net.out_len == 1000
    rng = np.random.RandomState(3)
    for mode in ("default", "strict"):
        set_mode(net, mode)
        RT, PT = MODES[mode]
        for N in ((3, 40) if mode == "default" else (3,)):
            img = (rng.rand(N, *chw) * 256 - 128).astype(np.float32)
            with ThreadPoolExecutor(8) as ex:
                ref = list(ex.map(lambda i: po.net_forward(layers, params, img[i:i + 1], keep=True), range(N)))
            ref_prob = np.stack([r[0][0] for r in ref])
            logits = torch.empty((N, 1000), dtype=torch.float32, device="cuda")
            if N == 3:      # every feature map of the un-fused run
                net.set_keep_maps(True)
                prob_d = net.forward(torch.from_numpy(img).cuda(), logits=logits)   # keep alive: the last map lives in it
                for l in range(len(layers) + 1):
                    fm = net.featmap(l, N)
                    want = np.stack([r[1][l][0] for r in ref])
                    e = close(fm.cpu().numpy().reshape(-1), want.reshape(-1))
                    assert e <= RT, (name, mode, l, e)
                del prob_d
                net.set_keep_maps(False)
            prob = net.forward(torch.from_numpy(img).cuda(), logits=logits).cpu().numpy()
            assert close(logits.cpu().numpy(), np.stack([r[1][len(layers) - 1][0].reshape(-1) for r in ref])) <= RT, (name, mode, N)
            assert np.abs(prob - ref_prob).max() <= PT, (name, mode, N)
    net.close()
