import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import esp_audio_libs_b200 as espb
L = espb.lib(); espb.set_device(0)
ns, ch, frames = 32, 8, 960000
cap = int(frames * 44100 / 96000) + 64
r = espb.Resampler(ns, frames * ch, cap * ch, 96000, 44100, 24, 24, ch, True, True, 1024, 256)
if len(sys.argv) > 1:
    r.set_biquad_time_blocks(8192, 1024)
raw = np.random.default_rng(0).integers(0, 256, size=(ns, frames * ch * 3), dtype=np.uint8)
d_in = espb.DeviceBuffer.from_numpy(raw); out_row = (cap * ch * 3 + 15) & ~15; d_out = espb.DeviceBuffer(ns * out_row)
for it in range(4):
    L.espb_device_sync(); t0 = time.perf_counter()
    res = r.resample_dev(d_in.ptr, raw.shape[1], d_out.ptr, out_row, frames, cap, 0.0)
    t1 = time.perf_counter()
    print(it, "call %.2f ms" % ((t1 - t0) * 1e3), res["frames_generated"], flush=True)
