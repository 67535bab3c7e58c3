#!/usr/bin/env python
"""Clock groups (ASRC): 4096 stereo streams, 44.1 -> ~48 kHz at the C2 filter geometry, split into
1 / 64 / 1024 groups that each follow their own clock (a slightly different, per-call drifting ratio), through
espb_resampleGroupsProcessInterleaved with device buffers.  Steady-state streaming calls of 1 s and of 10 ms.
One JSON line per case: device time per call (CUDA events on the caller's stream), host time per call, G samples/s
and the fraction of the single-group rate.

    python tools/bench_groups.py
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import esp_audio_libs_b200 as espb  # noqa: E402

f32 = np.float32
L = espb.lib()


def run(total, n_groups, frames, calls, warm):
    ch, taps = 2, 256
    sizes = [total // n_groups] * n_groups
    g = espb.ResampleGroups(sizes, ch, taps, 256, 1.0, 3)
    for k in range(n_groups):
        g.advance(k, taps / 2)
    base = f32(48000) / f32(44100)
    cap = int(frames * float(base) * 1.01) + 16
    x = np.random.default_rng(1).uniform(-0.5, 0.5, (min(total, 256), frames * ch)).astype(f32)
    x = np.tile(x, (total // x.shape[0], 1))
    d_in, d_out = espb.DeviceBuffer.from_numpy(x), espb.DeviceBuffer(total * cap * ch * 4)
    n_in = (C.c_int * n_groups)(*([frames] * n_groups))
    n_out = (C.c_int * n_groups)(*([cap] * n_groups))
    res = (espb.capi._Result * n_groups)()
    stream = L.espb_stream_create()
    gen_total = [0]

    def call(i):
        # every group on its own clock, drifting from call to call
        ratios = (C.c_float * n_groups)(*[float(base * f32(1.0 + 1e-6 * (k + 1) + 1e-7 * i)) for k in range(n_groups)])
        rc = L.espb_resampleGroupsProcessInterleaved(g.h, d_in.ptr, frames * ch, n_in, d_out.ptr, cap * ch, n_out,
                                                     ratios, res, stream)
        assert rc == 0, espb.capi._err()
        gen_total[0] = sum(res[k].output_generated * sizes[k] for k in range(n_groups))

    for i in range(warm):
        call(i)
    L.espb_stream_sync(stream)
    ev0, ev1 = L.espb_event_create(), L.espb_event_create()
    L.espb_event_record(ev0, stream)
    t0 = time.perf_counter()
    for i in range(calls):
        call(warm + i)
    host = time.perf_counter() - t0
    L.espb_event_record(ev1, stream)
    ms = C.c_float(0)
    L.espb_event_elapsed_ms(ev0, ev1, C.byref(ms))
    per_call = ms.value / calls
    samples = gen_total[0] * ch
    out = dict(streams=total, groups=n_groups, streams_per_group=total // n_groups, frames_per_call=frames,
               device_ms_per_call=per_call, host_ms_per_call=host / calls * 1e3,
               gsamples_per_s=samples / per_call / 1e6)
    g.free()
    d_in.free()
    d_out.free()
    L.espb_stream_destroy(stream)
    return out


def main():
    espb.set_device(0)
    for frames, calls, warm in ((44100, 6, 3), (441, 100, 20)):
        single = None
        for groups in (1, 64, 1024):
            r = run(4096, groups, frames, calls if groups < 1024 else max(3, calls // 4), warm)
            if groups == 1:
                single = r["gsamples_per_s"]
            r["fraction_of_single_group_rate"] = r["gsamples_per_s"] / single
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
