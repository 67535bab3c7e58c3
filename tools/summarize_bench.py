#!/usr/bin/env python
"""One line per bench JSON: headline, e2e and the sub-records.  usage: tools/summarize_bench.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    d = json.load(open(f))
    e = d.get("e2e") or {}
    lp = e.get("link_probe") or {}
    print(f"{f}: N={d['n_gpus']} value {d['value'] / 1e3:.1f} G/s  {d['ms_per_step']:.3f} ms  kernel frac "
          f"{d['roofline']['frac']:.3f} (step-level {d['roofline'].get('step_level_frac') or 0:.3f})  e2e "
          f"{e.get('value', 0) / 1e3:.1f} G/s {e.get('ms_per_step', 0):.1f} ms  link/rank {e.get('link_gbs_this_rank', 0):.1f} "
          f"GB/s  probe h2d {lp.get('h2d_gbs', 0):.1f} d2h {lp.get('d2h_gbs', 0):.1f} duplex {lp.get('duplex_sum_gbs', 0):.1f}  "
          f"frac_of_link {e.get('frac_of_link') or 0:.2f}")
    for k, v in (d.get("configs") or {}).items():
        ee = v.get("e2e") or {}
        print(f"    {k}: {v['value'] / 1e3:.2f} G/s  {v['ms_per_step']:.2f} ms  kernel frac {v['roofline']['frac']:.3f}  "
              f"by rank {[round(t, 2) for t in v.get('ms_per_step_by_rank', [])]}  e2e {ee.get('value', 0) / 1e3:.2f} G/s "
              f"{ee.get('ms_per_step', 0):.1f} ms frac_of_link {ee.get('frac_of_link') or 0:.2f}  "
              f"checksum {v.get('checksum_of_checksums', '')}  parity {v.get('parity_vs_oracle', '')}")
