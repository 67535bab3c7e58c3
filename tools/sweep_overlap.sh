#!/bin/bash
# Staging overlap on the BASELINE configs[1] workload: step time with the resampler launched as a programmatic
# dependent of the transposing kernel (ESPB_OVERLAP=1, default) against the serial order, per staging-grid size.
# usage: tools/sweep_overlap.sh "ESPB_OVERLAP=0" "ESPB_STAGE_CTAS=1" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-configs 2>gpurun_out/sweep_overlap.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$cfg', 'step_ms=%.3f serial_step_ms=%.3f kernel_ms=%.3f frac=%.4f step_frac=%.4f checksum=%s' % (d['ms_per_step'], r['step_ms_without_overlap'], r['kernel_ms'], r['frac'], r['step_level_frac'], d['checksums'][0]))" || tail -5 gpurun_out/sweep_overlap.err
done
