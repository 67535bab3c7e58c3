#!/bin/bash
# few-series kernel: ring stages x slice size (run on a GPU box)
for cfg in "3 21" "2 21" "3 10" "4 10" "2 10" "4 5" "3 5"; do
  set -- $cfg
  echo "== stages $1 slice_kb $2"
  ESPB_FS_STAGES=$1 ESPB_FS_SLICE_KB=$2 python tools/bench_few.py --quick --seconds 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'case' in d:
        print('  ', d['case'][:34].ljust(34), 'kernel_ms', round(d['kernel_ms_per_call'], 3), 'frac', round(d['frac_of_ffma2_probe'], 3))
"
done
