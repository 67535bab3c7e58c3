import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import esp_audio_libs_b200 as espb
L = espb.lib(); espb.set_device(0)
ns, frames, cap = 16384, 16000, 48064
r = espb.Resampler(ns, frames, cap, 16000, 48000, 16, 16, 1, True, True, 256, 256)
raw = np.random.default_rng(0).integers(0, 256, size=(ns, frames * 2), dtype=np.uint8) if len(sys.argv) > 1 else np.zeros((ns, frames * 2), np.uint8)
d_in = espb.DeviceBuffer.from_numpy(raw); d_out = espb.DeviceBuffer(ns * cap * 2)
for it in range(6):
    L.espb_device_sync(); t0 = time.perf_counter()
    res = r.resample_dev(d_in.ptr, frames * 2, d_out.ptr, cap * 2, frames, cap, 0.0)
    t1 = time.perf_counter(); L.espb_device_sync(); t2 = time.perf_counter()
    print(it, "call %.2f ms, +sync %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), res["frames_generated"], flush=True)
