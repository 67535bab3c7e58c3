import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import esp_audio_libs_b200 as espb
L = espb.lib()
espb.set_device(0)
n = 1 << 30
h = espb.PinnedBuffer(n, np.uint8); h.array[:] = 1
h2 = espb.PinnedBuffer(n, np.uint8)
d = espb.DeviceBuffer(n); d2 = espb.DeviceBuffer(n)
s1, s2 = L.espb_stream_create(), L.espb_stream_create()
for name in ("h2d", "d2h", "both"):
    L.espb_device_sync()
    t0 = time.perf_counter()
    for _ in range(4):
        if name in ("h2d", "both"): L.espb_memcpy_h2d(d.ptr, h.ptr, n, s1)
        if name in ("d2h", "both"): L.espb_memcpy_d2h(h2.ptr, d2.ptr, n, s2)
    L.espb_stream_sync(s1); L.espb_stream_sync(s2)
    dt = time.perf_counter() - t0
    print(name, "GB/s per direction: %.1f" % (4 * n / dt / 1e9))
