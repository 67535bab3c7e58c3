#!/bin/bash
# End-to-end (host buffers) variants on the BASELINE configs[1] workload: ms per step per setting.
for cfg in "$@"; do
  env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$cfg', 'e2e_ms=%.2f e2e_Gs=%.2f' % (e['ms_per_step'], e['value']/1e3))"
done
