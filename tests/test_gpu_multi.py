"""Multi-GPU plumbing in C: one process driving several devices — one batch context per device on a
contiguous stream range, NO collective on the data path, one ncclAllGather of {checksum, frames, first stream} per
shard through espb_multi_* (ncclCommInitAll).  Needs >= 2 GPUs in the box; skipped otherwise.  The sharded run must
give, stream for stream, the bytes of the single-device run, and the gathered checksums must add up to its checksum."""
import ctypes as C

import numpy as np
import pytest
from conftest import bits_equal
from oracle_lib import noise

import esp_audio_libs_b200 as espb

pytestmark = pytest.mark.gpu
f32 = np.float32


def _need(n):
    if espb.device_count() < n:
        pytest.skip(f"needs {n} GPUs, this box has {espb.device_count()}")


def test_nccl_is_resolved():
    assert espb.lib().espb_nccl_version() >= 22000


def test_single_process_two_devices_sharded_batch_and_nccl_gather():
    _need(2)
    L = espb.lib()
    world = 2
    ns, ch, taps, n_in = 300, 2, 256, 3000   # 300 streams -> shards of 150 (partial series groups on both devices)
    ratio = f32(48000) / f32(44100)
    cap = int(n_in * float(ratio)) + 40
    x = np.stack([noise(n_in, ch, stream=s, amp=0.8) for s in range(ns)])

    def run(device, rows):
        espb.set_device(device)
        b = espb.ResampleBatch(rows.shape[0], ch, taps, 256, 1.0, 3, mode=espb.MODE_EXACT)
        b.advance(taps / 2)
        d_in, d_out = espb.DeviceBuffer.from_numpy(rows), espb.DeviceBuffer(rows.shape[0] * cap * ch * 4)
        d_out.zero()
        used, gen = b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio)
        y = d_out.download(f32).reshape(rows.shape[0], cap * ch)
        cs = espb.checksum_u32(d_out.ptr, rows.shape[0] * cap * ch)  # wrapping 64-bit sum of the output words
        b.free()
        return y, gen, cs

    y_all, gen_all, cs_all = run(0, x)
    m = L.espb_multi_create(world, None)
    assert m, L.espb_multi_last_error()
    assert L.espb_multi_size(m) == world
    words = (C.c_uint64 * (3 * world))()
    parts = []
    for r in range(world):
        first, count = C.c_int64(0), C.c_int64(0)
        L.espb_shard_range(ns, r, world, C.byref(first), C.byref(count))
        assert (first.value, count.value) == espb.shard_range(ns, r, world)
        y, gen, cs = run(L.espb_multi_device(m, r), x[first.value:first.value + count.value])
        parts.append(y)
        words[3 * r:3 * r + 3] = [cs, gen, first.value]
    espb.set_device(0)
    gathered = (C.c_uint64 * (3 * world))()
    assert L.espb_multi_gather_words(m, words, 3, gathered) == 0, L.espb_multi_last_error()
    assert list(gathered) == list(words)           # rank order, unchanged by the trip over NCCL
    assert bits_equal(np.concatenate(parts), y_all)  # shard for shard the single-device bytes
    assert all(gathered[3 * r + 1] == gen_all for r in range(world))
    total = espb.combine_checksums([gathered[3 * r] for r in range(world)])
    assert total == cs_all % (1 << 64)  # order-independent: the shards' sums add up to the whole batch's
    L.espb_multi_free(m)


def test_host_link_probe_runs_on_all_devices():
    n = espb.device_count()
    r = espb.measure_host_link(list(range(min(n, 2))), 64 << 20, 16 << 20, 1)
    assert r["h2d_gbs"] > 1.0 and r["d2h_gbs"] > 1.0 and r["duplex_sum_gbs"] > 1.0
