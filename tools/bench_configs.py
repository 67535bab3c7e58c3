#!/usr/bin/env python
"""Times the other BASELINE.json configurations (C3, C4, C5 units) through the public C ABI on one GPU.
Not the bench contract (bench.py is); this produces the per-config numbers quoted in DESIGN.md.

    python tools/bench_configs.py [c3] [c4] [c5] [--scale S]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import esp_audio_libs_b200 as espb  # noqa: E402

f32 = np.float32
L = espb.lib()


def timed(fn, reps=7, warm=3):
    """Median per-call time from CUDA events around each call (first calls allocate scratch)."""
    for _ in range(warm):
        fn()
    L.espb_device_sync()
    times = []
    for _ in range(reps):
        ev0, ev1 = L.espb_event_create(), L.espb_event_create()
        L.espb_event_record(ev0, None)
        fn()
        L.espb_event_record(ev1, None)
        ms = espb.capi.C.c_float(0)
        L.espb_event_elapsed_ms(ev0, ev1, espb.capi.C.byref(ms))
        times.append(ms.value)
        L.espb_event_destroy(ev0)
        L.espb_event_destroy(ev1)
    times.sort()
    return times[len(times) // 2], times


def pcm_noise(ns, n, bits, rng):
    nb = (bits + 7) // 8
    base = (rng.normal(0, 0.25, size=(min(ns, 64), n)).clip(-1, 0.9999) * (2 ** (8 * nb - 1))).astype(np.int64)
    raw = np.zeros((base.shape[0], n * nb), np.uint8)
    for b in range(nb):
        raw[:, b::nb] = (base >> (8 * b)) & 0xFF
    reps = (ns + raw.shape[0] - 1) // raw.shape[0]
    return np.tile(raw, (reps, 1))[:ns]


def run_wrapper(name, ns, ch, src, dst, sb, db, taps, filters, frames, mode=espb.MODE_FAST, blocks=None):
    rng = np.random.default_rng(1)
    cap = int(frames * dst / src) + 64
    r = espb.Resampler(ns, frames * ch, cap * ch, src, dst, sb, db, ch, True, True, taps, filters, mode=mode)
    if blocks:
        r.set_biquad_time_blocks(*blocks)
    raw = pcm_noise(ns, frames * ch, sb, rng)
    d_in = espb.DeviceBuffer.from_numpy(raw)
    ob = (db + 7) // 8
    out_row = (cap * ch * ob + 15) & ~15
    d_out = espb.DeviceBuffer(ns * out_row)
    res = {}

    def step():
        # steady-state streaming: state carries from call to call (no reset), as a real stream would
        res["r"] = r.resample_dev(d_in.ptr, raw.shape[1], d_out.ptr, out_row, frames, cap, 0.0)

    ms, all_ms = timed(step)
    gen = res["r"]["frames_generated"]
    # the same calls enqueued without synchronising (espb_resampler_resample_async): host planning of call k+1
    # overlaps the device work of call k
    n_async = 8
    L.espb_device_sync()
    ev0, ev1 = L.espb_event_create(), L.espb_event_create()
    L.espb_event_record(ev0, None)
    for _ in range(n_async):
        r.resample_dev_async(d_in.ptr, raw.shape[1], d_out.ptr, out_row, frames, cap, 0.0)
    L.espb_event_record(ev1, None)
    ms_a = espb.capi.C.c_float(0)
    L.espb_event_elapsed_ms(ev0, ev1, espb.capi.C.byref(ms_a))
    ms_async = ms_a.value / n_async
    samples = gen * ch * ns
    pol = r.policy()
    line = dict(config=name, streams=ns, channels=ch, src_rate=src, dst_rate=dst, bits=(sb, db), taps=taps,
                filters=filters, frames_in=frames, frames_out=gen, filter=pol["filter"], ms_per_call=ms,
                ms_all=[round(t, 2) for t in all_ms], ms_per_call_pipelined=ms_async,
                msamples_per_s=samples / ms / 1e3, flop_per_sample=4 * taps,
                resampler_tflops_if_all_time=4 * taps * samples / ms / 1e9)
    print(json.dumps(line), flush=True)
    r.free()


def run_streaming(ns, ch, frames, calls=200):
    """Real-time use: small chunks (10 ms) through the float resampler, state carried on the device from call to
    call; every call has a new fractional position, so nothing is reused from the plan cache.  Host wall clock per
    call (schedule + uploads + launches) next to the device time."""
    taps = filters = 256
    ratio = f32(48000) / f32(44100)
    b = espb.ResampleBatch(ns, ch, taps, filters, 1.0, 3)
    b.advance(taps / 2)
    cap = int(frames * float(ratio)) + 8
    x = np.random.default_rng(2).uniform(-0.5, 0.5, (ns, frames * ch)).astype(f32)
    d_in, d_out = espb.DeviceBuffer.from_numpy(x), espb.DeviceBuffer(ns * cap * ch * 4)
    for _ in range(20):
        b.process_interleaved_dev(d_in.ptr, frames * ch, frames, d_out.ptr, cap * ch, cap, ratio)
    L.espb_device_sync()
    ev0, ev1 = L.espb_event_create(), L.espb_event_create()
    L.espb_event_record(ev0, None)
    t0 = time.perf_counter()
    gen = 0
    for _ in range(calls):
        gen += b.process_interleaved_dev(d_in.ptr, frames * ch, frames, d_out.ptr, cap * ch, cap, ratio)[1]
    t_host = time.perf_counter() - t0
    L.espb_event_record(ev1, None)
    ms = espb.capi.C.c_float(0)
    L.espb_event_elapsed_ms(ev0, ev1, espb.capi.C.byref(ms))
    print(json.dumps(dict(config=f"streaming {ns} x {ch}ch float, 44.1->48 kHz, {frames}-frame chunks", calls=calls,
                          host_us_per_call=t_host / calls * 1e6, device_us_per_call=ms.value / calls * 1e3,
                          msamples_per_s=gen * ch * ns / ms.value / 1e3,
                          realtime_factor=(frames / 44100.0) / (ms.value / calls / 1e3))), flush=True)
    b.free()


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c3", "c4", "c5"]
    if "stream" in which:
        espb.set_device(0)
        run_streaming(1024, 2, 441)
        run_streaming(4096, 2, 441)
        run_streaming(64, 2, 441)
    scale = 1.0
    if "--scale" in sys.argv:
        scale = float(sys.argv[sys.argv.index("--scale") + 1])
    espb.set_device(0)
    if "c3" in which:  # 16 kHz -> 48 kHz mono voice, post biquad, int16 in/out, 16384 streams, 1 s per call
        run_wrapper("C3 16k->48k mono int16, post-biquad", int(16384 * scale), 1, 16000, 48000, 16, 16, 256, 256, 16000)
    if "c4" in which:  # 96 kHz -> 44.1 kHz, 8 channels, 24-bit, 1024 taps, long streams (10 s per call), few streams
        run_wrapper("C4 96k->44.1k 8ch int24 T=1024, pre-biquad", max(int(32 * scale), 1), 8, 96000, 44100, 24, 24,
                    1024, 256, 960000)
        run_wrapper("C4 same, pre-biquad in time blocks of 8192 frames (1024 warm-up)", max(int(32 * scale), 1), 8,
                    96000, 44100, 24, 24, 1024, 256, 960000, blocks=(8192, 1024))
    if "c5" in which:  # 48 -> 44.1 kHz stereo f32-equivalent (32-bit PCM), low-pass, 8192 streams (= 65536 / 8 GPUs)
        run_wrapper("C5 shard 48k->44.1k stereo int32, pre-biquad, 8192 streams", int(8192 * scale), 2, 48000, 44100,
                    32, 32, 256, 256, 48000)


if __name__ == "__main__":
    main()
