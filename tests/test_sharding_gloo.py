"""Multi-GPU host logic on the CPU: world_size 2 over gloo.  Each rank takes its contiguous stream
range, produces per-stream results (here with the oracle, which is only the stand-in for the device
work in this CPU test), and the ranks all-gather checksum + counts exactly as bench.py does."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

import esp_audio_libs_b200 as espb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            got = [espb.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            sizes = [c for _, c in got]
            assert max(sizes) - min(sizes) <= 1
    assert espb.shard_range(65536, 3, 8) == (3 * 8192, 8192)  # 8192 stereo streams per GPU


def test_combine_checksums_wraps():
    assert espb.combine_checksums([2 ** 64 - 1, 2]) == 1
    assert espb.combine_checksums([]) == 0


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch.distributed as dist
    import esp_audio_libs_b200 as espb
    from oracle_lib import Oracle, noise
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_streams, ch, taps, n_in = 10, 2, 32, 400
    ratio = np.float32(48000) / np.float32(44100)
    first, count = espb.shard_range(n_streams, rank, world)
    o = Oracle()
    csum, frames = 0, 0
    for s in range(first, first + count):
        ctx = o.resampler(ch, taps, 32, 1.0, 3)
        y, used, gen = ctx.process_interleaved(noise(n_in, ch, stream=s), 600, ratio)
        csum = (csum + int(y.view(np.uint32).astype(np.uint64).sum())) & 0x7FFFFFFFFFFFFFFF
        frames += gen
    rows = espb.gather_words([csum, frames, count], dist)
    if rank == 0:
        print("RESULT", rows)
    dist.destroy_process_group()
""")


def test_two_rank_gather_over_gloo(tmp_path, oracle):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT")][0]
    rows = eval(line[len("RESULT "):])
    assert len(rows) == 2 and [r[2] for r in rows] == [5, 5]
    # the same batch in one process
    from oracle_lib import noise
    ratio = np.float32(48000) / np.float32(44100)
    total, frames = 0, 0
    for s in range(10):
        ctx = oracle.resampler(2, 32, 32, 1.0, 3)
        y, _, gen = ctx.process_interleaved(noise(400, 2, stream=s), 600, ratio)
        total = (total + int(y.view(np.uint32).astype(np.uint64).sum())) & 0x7FFFFFFFFFFFFFFF
        frames += gen
    assert (rows[0][0] + rows[1][0]) & 0x7FFFFFFFFFFFFFFF == total
    assert rows[0][1] + rows[1][1] == frames
