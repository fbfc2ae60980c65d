"""Batch sharding of the PQ forward pass across ranks: every image is independent, weights are
replicated, rank r owns a contiguous slice of the batch, and the only exchange is one all-gather of the per-rank
output rows ([rows, 1000] logits or probabilities).  Works with NCCL (GPU tensors, all_gather_into_tensor over
NVLink) and gloo (CPU tensors; used by the world_size-2 CPU tests)."""


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`: floor split, remainder spread over the first ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total, group=None, out=None):
    """Gathers per-rank row blocks (possibly of unequal height) into the full [n_total, C] tensor on every rank.
    `out` (equal shards, NCCL): a caller-owned [n_total, C] buffer to gather into (no allocation inside the step)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    cols = local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    heights = [hi - lo for lo, hi in sizes]
    if len(set(heights)) == 1 and dist.get_backend(group) == "nccl":
        if out is None:
            out = torch.empty((n_total, cols), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    hmax = max(heights)
    pad = torch.zeros((hmax, cols), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:h] for p, h in zip(parts, heights)], dim=0)


class OverlappedGather(object):
    """The path's only exchange, taken off the critical path: the all-gather of step i's rows runs on a side stream
    (it waits for step i's last kernel through an event) while the compute stream starts step i+1.  Two buffer sets
    alternate, so the rows being gathered are never overwritten; `rows(i)` is where step i must write its output."""

    def __init__(self, rows_per_rank, cols, world, device):
        import torch
        self.world, self.n_total = world, rows_per_rank * world
        self.local = [torch.empty((rows_per_rank, cols), dtype=torch.float32, device=device) for _ in range(2)]
        self.full = [torch.empty((self.n_total, cols), dtype=torch.float32, device=device) for _ in range(2)]
        self.side = torch.cuda.Stream(device=device)
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.steps = 0

    def rows(self, i):
        import torch
        b = i & 1
        if self.steps >= 2:      # the gather that last read this buffer set (two steps ago) must be over
            torch.cuda.current_stream().wait_event(self.done[b])
        return self.local[b]

    def launch(self, i):
        """Call after step i's kernels were enqueued on the current stream; returns the buffer the result lands in."""
        import torch
        b = i & 1
        self.ready[b].record(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready[b])
            all_gather_rows(self.local[b], self.n_total, out=self.full[b])
            self.done[b].record(self.side)
        self.steps += 1
        return self.full[b]
