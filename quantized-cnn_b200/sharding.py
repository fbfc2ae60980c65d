"""Batch sharding of the PQ forward pass across ranks: every image is independent, weights are
replicated, rank r owns a contiguous slice of the batch, and the only exchange is one all-gather of the per-rank
output rows ([rows, 1000] logits or probabilities).  Works with NCCL (GPU tensors, all_gather_into_tensor over
NVLink) and gloo (CPU tensors; used by the world_size-2 CPU tests)."""


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of rank `rank`: floor split, remainder spread over the first ranks."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total, group=None):
    """Gathers per-rank row blocks (possibly of unequal height) into the full [n_total, C] tensor on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    cols = local.shape[1]
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    heights = [hi - lo for lo, hi in sizes]
    if len(set(heights)) == 1 and dist.get_backend(group) == "nccl":
        out = torch.empty((n_total, cols), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    hmax = max(heights)
    pad = torch.zeros((hmax, cols), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:h] for p, h in zip(parts, heights)], dim=0)
