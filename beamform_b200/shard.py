"""Multi-GPU partitioning of the hot path: independent streams are split into contiguous
blocks, one block per rank, with no data-path collective.  The only exchange is the gather of steered-response maps."""


def shard_streams(n_streams, world_size, rank):
    """Contiguous block [begin, end) of rank `rank`: sizes differ by at most one, earlier ranks take the extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(n_streams, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_maps(local_maps, n_streams, group=None):
    """All-gather steered-response maps [b_local][T][D] -> [n_streams][T][D] on every rank (torch.distributed; NCCL over
    NVLink on GPUs, gloo in the CPU tests).  Uneven shards are padded to the largest block for the collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_streams(n_streams, world, r) for r in range(world)]
    bmax = max(e - b for b, e in sizes)
    pad = torch.zeros((bmax,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    pad[: local_maps.shape[0]] = local_maps
    out = torch.empty((world * bmax,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * bmax: r * bmax + (e - b)] for r, (b, e) in enumerate(sizes)], dim=0)
