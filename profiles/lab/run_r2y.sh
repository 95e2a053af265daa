#!/bin/bash
# round-2 lab y: reference (level-last) layout read in place by the sorted CAPE execution vs relayout + level-major
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "layouts or full_field or sorted_execution or garbage or multi_entry or top_first or level_order or dtype or float64 or host" 2>&1 | tail -3
for direct in 1 0; do
  for wl in C2 C3 C5; do
    XCAPE_B200_CAPE_DIRECT=$direct python bench.py --workload $wl --no-extras --no-cpu-baseline --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('DIRECT=$direct $wl device path ms/step', round(r['ms_per_step'], 3), 'call(level-major) ms', round(r['roofline']['call_ms'], 3), 'e2e ms', round(r['e2e']['ms_per_step'], 2), 'pageable', round(r['e2e']['pageable']['ms_per_step'], 2))
"
  done
done
} > gpurun_out/r2y_lab.txt 2>&1
cat gpurun_out/r2y_lab.txt
