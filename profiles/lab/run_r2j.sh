#!/bin/bash
# round-2 lab j: sort-key variants of the sorted CAPE execution; SRH tile kernel with cp.async double buffering
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "srh or sorted_execution" 2>&1 | tail -3
for tb in 0.25 1 2 4 8; do
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_TBIN=$tb python profiles/lab_time_kernel.py C2 2 10
  XCAPE_B200_SORT=1 XCAPE_B200_SORT_TBIN=$tb python profiles/lab_time_kernel.py C3 3 5
done
XCAPE_B200_SORT=1 XCAPE_B200_SORT_TBIN=2 python profiles/lab_time_kernel.py C5 2 5
for tile in 1 0; do
  XCAPE_B200_SRH_TILE=$tile python bench.py --workload C4 --no-extras --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('TILE=$tile value', r['value'], 'ms/step', r['ms_per_step'], 'kernel_ms', r['roofline']['kernel_ms'], 'e2e', r['e2e']['value'])
"
done
ncu --set full --clock-control none --import-source on -k regex:srh_tile -c 1 -o gpurun_out/r2j_srh_tile python bench.py --workload C4 --no-extras --steps 1 --warmup 1 > gpurun_out/r2j_ncu.log 2>&1
ncu -i gpurun_out/r2j_srh_tile.ncu-rep --page raw --csv > gpurun_out/r2j_srh_tile_raw.csv 2>/dev/null
} > gpurun_out/r2j_lab.txt 2>&1
grep -v "^+" gpurun_out/r2j_lab.txt | tail -30
