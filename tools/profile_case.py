#!/usr/bin/env python
"""Small fixed workloads for ncu captures (a number printed by a run under ncu is never a bench value).
    python tools/profile_case.py c1     one stereo stream, 10 s, few-series kernel (3 calls)
    python tools/profile_case.py c4     one 8-channel stream, 96 -> 44.1 kHz, T = 1024, 5 s (2 calls)
    python tools/profile_case.py c2     4096 stereo streams, 1 s, standard kernel (2 calls)
    python tools/profile_case.py bq     stand-alone biquad, 16384 mono series x 32768 frames (2 calls)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import esp_audio_libs_b200 as espb  # noqa: E402

f32 = np.float32
case = sys.argv[1] if len(sys.argv) > 1 else "c1"
espb.set_device(0)
L = espb.lib()
if case == "bq":
    n = 16384 * 32768
    d = espb.DeviceBuffer(n * 4)
    d.zero()
    bq = espb.BiquadBatch(16384, 2, espb.biquad_lowpass(1.0 / 6.0))
    for _ in range(2):
        bq.apply_dev(d.ptr, (32768, 1, 1), 1, 32768)
    L.espb_device_sync()
    sys.exit(0)
ns, ch, taps, lp, flags, src, dst, sec, calls = {
    "c1": (1, 2, 256, 1.0, 3, 44100, 48000, 10, 3),
    "c4": (1, 8, 1024, float(f32(44100) / f32(96000) * (f32(1.0) - f32(10.24) / f32(1024))), 5, 96000, 44100, 5, 2),
    "c2": (4096, 2, 256, 1.0, 3, 44100, 48000, 1, 2),
}[case]
ratio = f32(dst) / f32(src)
n_in = src * sec
cap = int(n_in * float(ratio)) + 64
x = np.random.default_rng(1).uniform(-0.5, 0.5, (min(ns, 64), n_in * ch)).astype(f32)
x = np.tile(x, (ns // x.shape[0], 1))
d_in, d_out = espb.DeviceBuffer.from_numpy(x), espb.DeviceBuffer(ns * cap * ch * 4)
b = espb.ResampleBatch(ns, ch, taps, 256, lp, flags)
b.set_option(espb.OPT_PLAN_CACHE, 0)
for _ in range(calls):
    b.reset()
    b.advance(taps / 2)
    print(b.process_interleaved_dev(d_in.ptr, n_in * ch, n_in, d_out.ptr, cap * ch, cap, ratio))
L.espb_device_sync()
