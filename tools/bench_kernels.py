#!/usr/bin/env python
"""HBM-bound kernels of the path on their own, through the public C ABI on one GPU: achieved GB/s of the algorithmic
bytes against the measured copy bandwidth in MEASURED_PEAKS.json.  Not the bench contract (bench.py is);
this produces the per-kernel numbers quoted in DESIGN.md §4.2.

    python tools/bench_kernels.py            -> one JSON line per kernel
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import esp_audio_libs_b200 as espb  # noqa: E402

C = espb.capi.C
L = espb.lib()
PEAK = 6500.3
try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
        PEAK = float(json.load(fh)["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    L.espb_device_sync()
    best = 1e30
    for _ in range(reps):
        ev0, ev1 = L.espb_event_create(), L.espb_event_create()
        L.espb_event_record(ev0, None)
        fn()
        L.espb_event_record(ev1, None)
        ms = C.c_float(0)
        L.espb_event_elapsed_ms(ev0, ev1, C.byref(ms))
        best = min(best, ms.value)
        L.espb_event_destroy(ev0)
        L.espb_event_destroy(ev1)
    return best


def report(name, nbytes, ms, note=""):
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(kernel=name, algorithmic_bytes=int(nbytes), ms=round(ms, 4), gb_per_s=round(gbs, 1),
                          frac_of_copy_peak=round(gbs / PEAK, 3), peak_gb_per_s=PEAK, note=note)), flush=True)


def main():
    espb.set_device(0)
    n = 1 << 29  # samples: 512 Mi (inputs and outputs far larger than the 126 MB L2)
    d_f = espb.DeviceBuffer(n * 4)
    d_b = espb.DeviceBuffer(n * 4)
    d_c = espb.DeviceBuffer(n * 4)
    d_f.zero()
    d_b.zero()
    clip = espb.DeviceBuffer(4)
    clip.zero()
    for bits, nb in ((16, 2), (24, 3), (32, 4)):
        ms = timed(lambda: L.espb_quantized_to_float(d_b.ptr, d_f.ptr, n, bits, 0.0, None))
        report(f"quantized_to_float {bits}-bit", n * (nb + 4), ms)
        ms = timed(lambda: L.espb_float_to_quantized(d_f.ptr, d_b.ptr, n, bits, clip.ptr, None))
        report(f"float_to_quantized {bits}-bit", n * (nb + 4), ms, "clip count accumulated on the device")
    # Q15 helpers
    ms = timed(lambda: L.espb_dsps_add_s16(d_f.ptr, d_b.ptr, d_c.ptr, 2 * n, 1, 1, 1, 1, None))
    report("dsps_add_s16", 2 * n * 6, ms)
    ms = timed(lambda: L.espb_dsps_mulc_s16(d_f.ptr, d_c.ptr, 2 * n, 23170, 1, 1, None))
    report("dsps_mulc_s16", 2 * n * 4, ms)
    # biquad pair, in place, many series (C3-like: 16384 mono series) and interleaved stereo
    for ch, streams, frames in ((1, 16384, 32768), (2, 4096, 65536)):
        bq = espb.BiquadBatch(streams * ch, 2, espb.biquad_lowpass(1.0 / 6.0))
        row = frames * ch
        ms = timed(lambda: bq.apply_dev(d_f.ptr, (row, 1, ch), ch, frames))
        report(f"biquad_apply_buffer x2 sections, {streams} streams x {ch} ch x {frames} frames (in place)",
               streams * row * 8, ms, "includes the layout stages to and from time-major rows")
        bq.free()
    # checksum (read only)
    acc = espb.DeviceBuffer(8)
    acc.zero()
    ms = timed(lambda: L.espb_checksum_u32(d_f.ptr, n, acc.ptr, None))
    report("checksum_u32", n * 4, ms)


if __name__ == "__main__":
    main()
