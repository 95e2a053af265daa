#!/bin/bash
# round-2 lab i: SRH tile kernel (reference layout read in place) — parity tests, A/B timing, ncu
set -x
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "srh" 2>&1 | tail -5
for tile in 1 0; do
  XCAPE_B200_SRH_TILE=$tile python bench.py --workload C4 --no-extras --steps 10 --warmup 3 2>&1 | tail -1 | python -c "
import sys, json
r = json.loads(sys.stdin.read())
print('TILE=$tile value', r['value'], 'ms/step', r['ms_per_step'], 'kernel_ms', r['roofline']['kernel_ms'], 'e2e', r['e2e']['value'])
print(json.dumps(r)[:1500])
"
done
ncu --set full --clock-control none --import-source on -k regex:srh_tile -c 1 -o gpurun_out/r2i_srh_tile python bench.py --workload C4 --no-extras --steps 1 --warmup 1 > gpurun_out/r2i_ncu.log 2>&1
ncu -i gpurun_out/r2i_srh_tile.ncu-rep --page raw --csv > gpurun_out/r2i_srh_tile_raw.csv 2>/dev/null
} > gpurun_out/r2i_lab.txt 2>&1
tail -30 gpurun_out/r2i_lab.txt
