#!/usr/bin/env python
"""Few streams per context (the reference's own use case): 1 / 4 / 16 / 64 stereo streams of 10 s at the C2 filter
geometry (44.1 -> 48 kHz, T = F = 256), and one 8-channel 96 -> 44.1 kHz stream at T = 1024 (C4's literal shape),
through the float C-ABI call with device buffers.  For each: the resampler kernel alone (CUDA events around its
launches: TFLOP/s against the FFMA2 probe of the same run) and the whole call, with the few-series kernel
(ESPB_FS unset) and through the standard kernel (ESPB_FS=0).  One JSON line per case.

    python tools/bench_few.py [--seconds 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import esp_audio_libs_b200 as espb  # noqa: E402

f32 = np.float32
L = espb.lib()


def run(ns, ch, taps, filters, lowpass, flags, src, dst, seconds, fs, mode, peak):
    if fs is None:
        os.environ.pop("ESPB_FS", None)
    else:
        os.environ["ESPB_FS"] = fs
    ratio = f32(dst) / f32(src)
    n_in = int(src * seconds)
    cap = int(n_in * float(ratio)) + 64
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, (ns, n_in * ch)).astype(f32)
    d_in, d_out = espb.DeviceBuffer.from_numpy(x), espb.DeviceBuffer(ns * cap * ch * 4)
    b = espb.ResampleBatch(ns, ch, taps, filters, lowpass, flags, mode=mode)
    b.set_option(espb.OPT_PLAN_CACHE, 0)
    b.set_option(espb.OPT_KERNEL_TIMING, 1)
    stream = L.espb_stream_create()

    def step():
        b.reset(stream)
        b.advance(taps / 2.0)
        return b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio, stream)

    for _ in range(3):
        used, gen = step()
    L.espb_stream_sync(stream)
    b.kernel_time()
    reps = 10
    ev0, ev1 = L.espb_event_create(), L.espb_event_create()
    L.espb_event_record(ev0, stream)
    for _ in range(reps):
        step()
    L.espb_event_record(ev1, stream)
    ms = espb.capi.C.c_float(0)
    L.espb_event_elapsed_ms(ev0, ev1, espb.capi.C.byref(ms))
    k_ms, k_n = b.kernel_time()
    samples = gen * ch * ns
    flop = 4.0 * taps * samples
    k_ms_call = k_ms / reps
    tf = flop / (k_ms_call * 1e-3) / 1e12
    line = dict(case=f"{ns} x {ch}ch {src}->{dst} T={taps} F={filters} {seconds}s", kernel="few-series" if fs is None else "standard (ESPB_FS=0)",
                mode="exact" if mode == espb.MODE_EXACT else "fast", frames_out=gen, series=ns * ch,
                kernel_ms_per_call=k_ms_call, kernel_launches_per_call=k_n / reps, kernel_tflops=tf,
                frac_of_ffma2_probe=tf / peak if mode != espb.MODE_EXACT else None, call_ms=ms.value / reps,
                call_msamples_per_s=samples / (ms.value / reps) / 1e3)
    print(json.dumps(line), flush=True)
    b.free()
    L.espb_stream_destroy(stream)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--quick", action="store_true", help="few-series kernel only: 1 and 4 stereo streams, 1 x 8 ch")
    args = ap.parse_args()
    espb.set_device(0)
    _, peak = espb.measure_fp32_fma_peak2()
    print(json.dumps(dict(ffma2_probe_tflops=peak)), flush=True)
    if args.quick:
        lp = float(f32(44100) / f32(96000) * (f32(1.0) - f32(10.24) / f32(1024)))
        run(1, 2, 256, 256, 1.0, 3, 44100, 48000, args.seconds, None, espb.MODE_FAST, peak)
        run(2, 2, 256, 256, 1.0, 3, 44100, 48000, args.seconds, None, espb.MODE_FAST, peak)
        run(4, 2, 256, 256, 1.0, 3, 44100, 48000, args.seconds, None, espb.MODE_FAST, peak)
        run(1, 8, 1024, 256, lp, 1 | 4, 96000, 44100, args.seconds, None, espb.MODE_FAST, peak)
        run(1, 2, 256, 256, float(f32(44100) / f32(48000) * f32(0.96)), 1, 48000, 44100, args.seconds, None,
            espb.MODE_FAST, peak)
        return
    for ns in (1, 2, 4, 16, 64):
        for fs in ((None, "0") if ns <= 16 else ("0",)):
            run(ns, 2, 256, 256, 1.0, 3, 44100, 48000, args.seconds, fs, espb.MODE_FAST, peak)
    run(1, 2, 256, 256, 1.0, 3, 44100, 48000, args.seconds, None, espb.MODE_EXACT, peak)
    lp = float(f32(44100) / f32(96000) * (f32(1.0) - f32(10.24) / f32(1024)))
    for fs in (None, "0"):
        run(1, 8, 1024, 256, lp, 1 | 4, 96000, 44100, args.seconds, fs, espb.MODE_FAST, peak)
    run(4, 8, 1024, 256, lp, 1 | 4, 96000, 44100, args.seconds, None, espb.MODE_FAST, peak)


if __name__ == "__main__":
    main()
