"""N > 1 path on CPU: world_size-2 gloo processes shard a batch, run the (oracle) forward on their slice and
all-gather the logits -- the same sharding/gather helper bench.py and a multi-GPU caller use with NCCL."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("quantized-cnn_b200.sharding")
    from oracle import pyoracle as po
    layers = [po.conv(1, 3, 16, 1, 1), po.relu(), po.pool(0, 2, 2), po.fcnt(16), po.relu(), po.drpt(0.5), po.fcnt(8),
              po.smax()]
    pq = {0: (1, 16, 4), 3: (36, 16, 4), 6: (4, 8, 4)}
    params = po.synth_model(layers, (4, 6, 6), pq, seed=5, ctrd_std=0.3)      # replicated weights (same seed)
    rng = np.random.RandomState(123)
    batch = (rng.randn(n_total, 4, 6, 6) * 3).astype(np.float32)              # every rank can index the full batch
    lo, hi = sh.shard_range(n_total, rank, world)
    local = po.net_forward(layers, params, batch[lo:hi]) if hi > lo else np.zeros((0, 8), np.float32)
    full = sh.all_gather_rows(torch.from_numpy(np.ascontiguousarray(local)), n_total)
    ref = po.net_forward(layers, params, batch)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.stack([full.numpy(), ref]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_two_rank_shard_and_gather(tmp_path, n_total):
    world = 2
    port = 29000 + (os.getpid() + n_total) % 2000
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        full, ref = np.load(os.path.join(str(tmp_path), "r%d.npy" % r))
        assert full.shape == (n_total, 8)
        assert np.array_equal(full, ref)      # sharded == unsharded, bit for bit, on every rank


def test_shard_ranges_cover_batch_exactly():
    sh = importlib.import_module("quantized-cnn_b200.sharding")
    for n in (1, 7, 8, 256, 8192, 8195):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
