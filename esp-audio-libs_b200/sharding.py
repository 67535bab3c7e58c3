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


class NcclGather:
    """One process per GPU (the torchrun-style launch of bench.py): the checksum gather and the barriers go through
    the library's own C entry points (espb_dist_*: ncclCommInitRank + ncclAllGather), not through torch.

    The only thing the launcher has to carry is NCCL's 128-byte unique id from rank 0 to the other ranks.  All ranks
    of the contract run on ONE node as children of one launcher process, so the id travels through a file named after
    that parent process and the rendezvous port (written atomically by rank 0, polled by the others)."""

    def __init__(self, rank, world, timeout_s=120.0):
        import ctypes as C
        import os
        import time

        from . import capi
        self._C, self._capi = C, capi
        self.rank, self.world = int(rank), int(world)
        L = capi.lib()
        if L.espb_nccl_version() <= 0:
            raise RuntimeError("NCCL (libnccl.so.2) not available")
        path = os.path.join(os.environ.get("ESPB_RENDEZVOUS_DIR", "/tmp"),
                            f"espb_nccl_{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}.id")
        buf = (C.c_char * 128)()
        if self.rank == 0:
            capi._check(L.espb_dist_unique_id(buf), "espb_dist_unique_id")
            with open(path + ".tmp", "wb") as fh:
                fh.write(bytes(buf))
            os.replace(path + ".tmp", path)
        else:
            t0 = time.time()
            while not os.path.exists(path):
                if time.time() - t0 > timeout_s:
                    raise RuntimeError(f"rank {rank}: no NCCL id at {path}")
                time.sleep(0.01)
            with open(path, "rb") as fh:
                data = fh.read()
            C.memmove(buf, data, 128)
        self.h = L.espb_dist_init(buf, self.rank, self.world)
        if not self.h:
            raise RuntimeError("espb_dist_init: " + (L.espb_multi_last_error() or b"").decode())
        self.barrier()
        if self.rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass

    def allgather(self, words):
        C = self._C
        words = [int(w) & 0xFFFFFFFFFFFFFFFF for w in words]
        n = len(words)
        mine = (C.c_uint64 * n)(*words)
        out = (C.c_uint64 * (n * self.world))()
        self._capi._check(self._capi.lib().espb_dist_allgather_u64(self.h, mine, n, out), "espb_dist_allgather_u64")
        return [[int(out[r * n + k]) for k in range(n)] for r in range(self.world)]

    def barrier(self):
        self._capi._check(self._capi.lib().espb_dist_barrier(self.h), "espb_dist_barrier")

    def max_float(self, value):
        """max over ranks of a non-negative float (gathered as its bit pattern)."""
        import struct
        bits = struct.unpack("<Q", struct.pack("<d", float(value)))[0]
        return max(struct.unpack("<d", struct.pack("<Q", g[0]))[0] for g in self.allgather([bits]))

    def close(self):
        if getattr(self, "h", None):
            self._capi.lib().espb_dist_free(self.h)
            self.h = None
