#!/usr/bin/env python
"""Write-only and read-only HBM bandwidth next to the copy peak: what a write-heavy streaming kernel (PCM -> float
writes 2 bytes for every byte it reads) can expect.  cudaMemsetAsync on 2 GiB / the library's checksum kernel."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import esp_audio_libs_b200 as espb  # noqa: E402

C = espb.capi.C
L = espb.lib()
espb.set_device(0)
n = 1 << 31
d = espb.DeviceBuffer(n)


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


ms = timed(lambda: L.espb_memset(d.ptr, 0, n, None))
print(json.dumps({"probe": "cudaMemsetAsync 2 GiB (write only)", "ms": round(ms, 4), "gb_per_s": round(n / ms / 1e6, 1)}))
