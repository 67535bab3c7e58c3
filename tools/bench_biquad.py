import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
import esp_audio_libs_b200 as espb
L=espb.lib()
espb.set_device(0)
def timed(fn, reps=5):
    for _ in range(2): fn()
    L.espb_device_sync()
    ts=[]
    for _ in range(reps):
        e0,e1=L.espb_event_create(),L.espb_event_create()
        L.espb_event_record(e0,None); fn(); L.espb_event_record(e1,None)
        ms=espb.capi.C.c_float(0); L.espb_event_elapsed_ms(e0,e1,espb.capi.C.byref(ms)); ts.append(ms.value)
    return sorted(ts)[len(ts)//2]
n=1<<29
d=espb.DeviceBuffer(n*4)
x=np.random.default_rng(0).uniform(-0.5,0.5,1<<24).astype(np.float32)
for k in range(0,n,1<<24):
    L.espb_memcpy_h2d(d.ptr+k*4, x.ctypes.data, (1<<24)*4, None)
L.espb_device_sync()
for onepass in ("1","0"):
    os.environ["ESPB_BIQUAD_ONEPASS"]=onepass
    for ch,streams,frames in ((1,16384,32768),(2,4096,65536),(4,2048,65536),(8,1024,65536)):
        bq=espb.BiquadBatch(streams*ch,2,espb.biquad_lowpass(1.0/6.0))
        row=frames*ch
        ms=timed(lambda: bq.apply_dev(d.ptr,(row,1,ch),ch,frames))
        print("onepass",onepass,ch,streams,frames,"ms",round(ms,3),"frac",round(streams*row*8/ms/1e6/6500.3,3),flush=True)
        bq.free()
