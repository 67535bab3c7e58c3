#!/usr/bin/env python
"""Host <-> device link ceiling for the end-to-end (host-buffer) entry points: plain pinned cudaMemcpyAsync, one call
per 64 MiB slab, H2D alone / D2H alone / both at once, on 1, 2, 4, 8 GPUs of the box concurrently (one process).
Writes profiles/r02_host_link.json.

    python tools/host_link_probe.py [--out profiles/r02_host_link.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import esp_audio_libs_b200 as espb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_host_link.json"))
ap.add_argument("--mbytes", type=int, default=1024)
args = ap.parse_args()
n_dev = espb.device_count()
rows = []
for n in (1, 2, 4, 8):
    if n > n_dev:
        break
    r = espb.measure_host_link(list(range(n)), args.mbytes << 20, 64 << 20, 3)
    r["gpus"] = n
    rows.append(r)
    print(json.dumps(r), flush=True)
info = espb.device_info()
doc = {"what": "pinned cudaMemcpyAsync, one call per 64 MiB slab, all listed GPUs concurrently; aggregate GB/s per "
               "direction (duplex_each: each direction while the other runs)", "device": info["name"],
       "gpus_on_box": n_dev, "cpu_count": os.cpu_count(), "rows": rows}
os.makedirs(os.path.dirname(args.out), exist_ok=True)
with open(args.out, "w") as fh:
    json.dump(doc, fh, indent=1)
