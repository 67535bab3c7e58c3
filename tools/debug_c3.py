import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import esp_audio_libs_b200 as espb
L=espb.lib(); espb.set_device(0)
ns, ch, sr, dr, bits, taps, filters, frames = 16384, 1, 16000, 48000, 16, 256, 256, 16000
cap = frames*3+64
raw = np.random.default_rng(3).integers(0,256,size=(ns, frames*2),dtype=np.uint8)
r = espb.Resampler(ns, frames*ch, cap*ch, sr, dr, bits, bits, ch, True, True, taps, filters)
r.set_option(espb.OPT_PLAN_CACHE, 0)
if len(sys.argv)>1 and sys.argv[1]=="timing": r.set_option(espb.OPT_KERNEL_TIMING, 1)
in_row, out_row = raw.shape[1], (cap*ch*2+15)&~15
d_in, d_out = espb.DeviceBuffer.from_numpy(raw), espb.DeviceBuffer(ns*out_row)
stream = L.espb_stream_create()
for i in range(10):
    t0=time.perf_counter()
    res = r.resample_dev_async(d_in.ptr, in_row, d_out.ptr, out_row, frames, cap, 0.0, stream)
    t1=time.perf_counter()
    L.espb_stream_sync(stream)
    t2=time.perf_counter()
    print(i, 'enqueue ms', round((t1-t0)*1e3,2), 'sync ms', round((t2-t1)*1e3,2), res['frames_generated'], flush=True)
