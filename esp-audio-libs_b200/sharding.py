"""Stream sharding across the GPUs of one node.

Streams share no state (the reference has no globals; every context is independent), so a batch is
cut into contiguous stream-index ranges, one per rank, and no collective touches the data path.
The only exchange is an all-gather of a few 64-bit words per rank (checksum, frame and clip counts)
after processing — NCCL on GPUs, gloo in the CPU tests.
"""


def shard_range(n_streams, rank, world):
    """Contiguous range [first, first + count) of rank `rank`: sizes differ by at most one stream."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(n_streams), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def combine_checksums(values):
    """Order-independent checksum of checksums: wrapping 64-bit sum."""
    total = 0
    for v in values:
        total = (total + int(v)) & 0xFFFFFFFFFFFFFFFF
    return total


def gather_words(words, dist=None, device=None):
    """All-gather a short list of non-negative ints (< 2**63) from every rank; returns a list per rank.

    `dist` is torch.distributed (initialised) or None for a single process."""
    words = [int(w) for w in words]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [words]
    import torch
    mine = torch.tensor(words, dtype=torch.int64, device=device)
    out = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [[int(x) for x in t.tolist()] for t in out]
