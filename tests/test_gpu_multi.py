"""Multi-GPU behind the C ABI (qcnn_multi_*: one process, batch-sharded replicas, ncclAllGather of the probabilities).

Runs with however many GPUs the box has: R = 1 exercises the same code (NCCL clique of one), R >= 2 the real exchange
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).  Expected values: the single-GPU network of
this library on the same images (whose parity against the CPU oracle is the subject of test_gpu_net / test_gpu_parity_b256),
bit-identical where the shard sizes give every replica the plans of the single-GPU run, within the mode tolerance
otherwise (tilings are chosen per batch size)."""
import numpy as np
import pytest

from test_gpu_net import MODES


@pytest.fixture(scope="module")
def model(po, tmp_path_factory):
    if po.have_alexnet():
        return po.ALEXNET_DIR, po.ALEXNET_PFX
    d = str(tmp_path_factory.mktemp("synth_multi"))
    po.save_model(d, "synth", po.synth_alexnet(seed=1))
    return d, "synth"


@pytest.mark.gpu
def test_multi_forward_matches_single_gpu(po, qcnn, ctx, model):
    import torch
    R = min(torch.cuda.device_count(), 8)
    dirpath, pfx = model
    PT = MODES["default"][1]
    net = qcnn.Net(ctx, dirpath, pfx, "AlexNet")
    for devs in sorted({1, R}):
        m = qcnn.MultiNet(dirpath, pfx, "AlexNet", devices=list(range(devs)))
        assert m.nccl_version >= 20000 and m.out_len == 1000
        for N in (1, 7, 64):
            img = po.lcg_images(N, 4000 + N)
            want = net.forward(torch.from_numpy(img).cuda()).cpu().numpy()
            # (a) host-buffer call == ExecForwardPass(img, prob)
            got = m.forward_host(torch.from_numpy(img).pin_memory())
            assert got.shape == (N, 1000)
            assert np.abs(got - want).max() <= PT, (devs, N, np.abs(got - want).max())
            assert np.array_equal(got.argmax(1), want.argmax(1))
            # (b) device-resident asynchronous steps: every GPU ends up with the whole [N, 1000] result; two steps in
            #     flight use the two buffer sets
            shards = []
            for r in range(devs):
                lo, hi = m.shard(N, r)
                shards.append(torch.from_numpy(img[lo:hi]).to("cuda:%d" % r) if hi > lo else None)
            p1 = m.forward(shards, N)
            p2 = m.forward(shards, N)
            m.sync()
            for r in range(devs):
                for p in (p1, p2):
                    g = m.gathered(p[r], N, r).cpu().numpy()
                    assert np.array_equal(g, got), (devs, N, r)
        m.close()
    net.close()


@pytest.mark.gpu
def test_multi_errors(qcnn, tmp_path):
    with pytest.raises(qcnn.QcnnError):
        qcnn.MultiNet(str(tmp_path), "nothing", "AlexNet", devices=[0])
    with pytest.raises(qcnn.QcnnError):
        qcnn.MultiNet(str(tmp_path), "nothing", "AlexNet", devices=[99])
