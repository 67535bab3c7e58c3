#!/usr/bin/env python
"""Non-interpolating mode (flags without SUBSAMPLE_INTERPOLATE) on the configs[1] shapes: kernel time with the
dedicated kernel (default) and, with ESPB_NI=0, through the interpolating kernel with an idle second filter."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import esp_audio_libs_b200 as espb  # noqa: E402

f32 = np.float32
espb.set_device(0)
ns, ch, taps, filters, n_in = 4096, 2, 256, 256, 44100
ratio = f32(48000) / f32(44100)
cap = int(n_in * float(ratio)) + 64
x = np.random.default_rng(1).uniform(-0.5, 0.5, (64, n_in * ch)).astype(f32)
x = np.tile(x, (ns // 64, 1))
d_in, d_out = espb.DeviceBuffer.from_numpy(x), espb.DeviceBuffer(ns * cap * ch * 4)
for mode in (espb.MODE_FAST, espb.MODE_EXACT):
    b = espb.ResampleBatch(ns, ch, taps, filters, 1.0, espb.BLACKMAN_HARRIS, mode=mode)
    b.set_option(espb.OPT_KERNEL_TIMING, 1)
    for _ in range(3):
        b.reset()
        b.advance(taps / 2)
        used, gen = b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio)
    b.kernel_time()
    for _ in range(10):
        b.reset()
        b.advance(taps / 2)
        b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio)
    espb.lib().espb_device_sync()
    ms, n = b.kernel_time()
    k = ms / n
    samples = gen * ch * ns
    # whole call, device-resident, with and without the staging overlap (kernel timing off)
    L = espb.lib()
    b.set_option(espb.OPT_KERNEL_TIMING, 0)
    b.set_option(espb.OPT_PLAN_CACHE, 0)
    call_ms = {}
    for overlap in (1, 0):
        b.set_option(espb.OPT_OVERLAP_STAGING, overlap)
        for it in range(13):
            if it == 3:
                ev0, ev1 = L.espb_event_create(), L.espb_event_create()
                L.espb_event_record(ev0, None)
            b.reset()
            b.advance(taps / 2)
            b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio)
        L.espb_event_record(ev1, None)
        t = espb.capi.C.c_float(0)
        L.espb_event_elapsed_ms(ev0, ev1, espb.capi.C.byref(t))
        call_ms[overlap] = round(t.value / 10, 3)
    print(json.dumps(dict(mode="fast" if mode == espb.MODE_FAST else "exact", dedicated=os.environ.get("ESPB_NI", "1") != "0",
                          kernel_ms=round(k, 3), gsamples_per_s=round(samples / k / 1e6, 1),
                          tflops_at_2T=round(2 * taps * samples / k / 1e9, 1), call_ms_overlap=call_ms[1],
                          call_ms_serial=call_ms[0])), flush=True)
    b.free()
