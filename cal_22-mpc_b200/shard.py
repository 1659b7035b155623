"""Multi-GPU sharding of a block stream: blocks are independent, so GPU g of G takes the
contiguous range [g*ceil(N/G), min(N, (g+1)*ceil(N/G))) and the only exchange is one all-reduce (sum) of the
statistics vector at the end.  One process per GPU; the collective is torch.distributed's (NCCL on GPUs,
gloo in the CPU tests)."""


def shard_range(n_blocks, rank, world):
    per = (n_blocks + world - 1) // world
    lo = min(n_blocks, rank * per)
    hi = min(n_blocks, (rank + 1) * per)
    return lo, hi


def allreduce_stats(stats_words, group=None):
    """In-place sum of the int64 statistics vector (a torch tensor on the device the backend needs) over ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats_words, op=dist.ReduceOp.SUM, group=group)
    return stats_words
